// Multi-view depth projection of packed ragged point clusters -- fused sm_100a kernels (one image per CTA pass).
//
// Replaces, for one (cluster, view) per CTA, everything between the canonicalised cluster and the
// CLIP patch embedding in the reference:
//   rotate                        src/utils/mv_utils.py:173-201  (point_transform, torch bmm)
//   normalise / ceil / clip       src/utils/mv_utils.py:99-118   (points2grid part 1)
//   3-D grid scatter-max          src/utils/mv_utils.py:120-127  (torch_scatter reduce="max")
//   5x5 max-pool, 3x3 Gaussian,   src/utils/mv_utils.py:30-37    (GridToImage.forward)
//   depth max, /max, 1-x
//   bilinear (R-2)->224, uint8    src/vilgod/zero_shot_detector.py:405-409
//   ToTensor / Normalize          third_party/CLIP/clip/clip.py:79-86 (folded into the patch-embed
//                                 weights; the kernel emits the integer pixel 0..255 as bf16 / fp16)
//
// Data flow: points are read from HBM (L2 for views > 0), every intermediate (one depth slice of
// the grid, the running depth-max image, the horizontally interpolated rows) lives in shared
// memory, and the only HBM writes are the patch-major tile (100,352 B per image) and, on request,
// the uint8 image.  Algorithmic bytes per cluster: 12 N + V * 224*224*2.
//
// Two kernels, bit-identical results:
//   projection_fast_kernel  clusters up to 2048 points (every cluster of a Waymo-shaped frame) whose touched
//          region fits its bounding-box buffer (always at the reference's obj_ratio): persistent CTAs, the
//          running depth-max image in registers, stamped max-pool, band-limited Gaussian, background as bulk
//          stores from a constant patch row, walking bilinear emit.  R = 112: three 256-thread CTAs per SM;
//          R = 224: one 512-thread CTA per SM.  What it does not take goes onto a device-side list ...
//   projection_kernel       ... which the general kernel works off (list mode); it also serves other
//          obj_ratios and the raw-grid debug tap (one CTA per image).  Two ways through its per-slice stencil:
//     stamp  (N <= 1024): the quantised points are counting-sorted by depth slice; per occupied slice the
//            5x5 max-pool is applied AT SCATTER TIME (each point stamps its 5x5 footprint with
//            shared-memory atomicMax: max-pool of a sparse grid is the union of the footprints), and the
//            3x3 Gaussian + depth max run only over the slice's bounding box.
//     dense  (larger clusters, and whenever the raw grid is requested): scatter-max of the raw cells, then
//            a row-streaming separable 5x5 max + Gaussian with register rings.
// The emit (bilinear, uint8 quantisation, patch-major tile) is restricted to the output rows AND
// column groups whose four source pixels can differ from background; everything else is the background
// value (one constant at R = 112 / 224, else copied from a precomputed tile).
//
// General kernel: R = 112: grid slice + image in shared memory, two CTAs per SM.  R = 224: the slice alone
// is 196 KB, so the running image lives in a per-SM scratch in global memory (L2 resident), one CTA per SM.
//
// Numerics contract (tests/test_projection_gpu.py): occupancy masks and scatter winners bit-exact
// against the oracle; every fp32 operation up to the scatter is a single IEEE-rounded operation in
// the reference's order (explicit __f*_rn intrinsics: no FMA contraction), the bilinear stage uses
// exactly the FMA pattern torch-CPU executes, the Gaussian uses a row-major FMA chain (the
// reference's conv order is unspecified; 1e-5 contract).
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace vg {
namespace {

constexpr int D = 8;            // depth slices
constexpr int S = 224;          // output image size
constexpr int NT = 512;         // threads per CTA
constexpr int NW = NT / 32;     // 16 warps
constexpr int CH = 7;           // max rows per warp chunk in the dense stencil pass
constexpr int CAP = 1408;       // points whose quantised (x, y, slice, value) are cached in smem
constexpr int POOL = 65536 - CAP;    // pooled points per resident CTA (N <= 65,536 never recomputes)
constexpr int kSpillPerSm = 2;  // resident CTAs per SM (shared memory bound)
constexpr int STAMP_N = 2 * NT; // largest cluster that takes the stamp path (two points per thread)
constexpr int NG = S / 8;       // 8-pixel output groups per row (28)

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
    float *img_scratch;    // R = 224: [nsmid][Q*R] running depth-max image of the CTA on that SM
    int32_t spill_sms;     // SM ids covered by the spill pool (%nsmid)
    int32_t C, V;
    float rot[VG_MAX_VIEWS * 9];
    float gauss[9];
    float obj_ratio, depth_bias, one_plus_bias;
    int32_t rotate_mode, div_mode;
    op_t *tiles;
    uint8_t *u8;
    int32_t u8_first_only; // u8 is [C,S,S] and receives view 0 only
    int32_t *status;
    float *dbg_grid;
    float *dbg_dens;
    // fast path hand-over: images the fast kernel does not take (more than FAST_N points, or a touched
    // region beyond its shared-memory layout) are appended here and run by projection_kernel in list mode
    int32_t *defer;        // [0] = count, [1] = cursor of the list-mode kernel, [2] = image counter of the
                           // fast kernel's persistent CTAs, [3 + i] = image index
    int32_t block0;        // first image of this launch (fast kernel, launches above the list capacity)
    long long *trace;      // VG_PROJ_TRACE: clock64 stamps of thread 0 at the phase boundaries, [image][12]
    uint32_t bg_splat;     // two operand-typed background pixels when the whole background tile is one value
                           // (R = 112: 255 everywhere), else 0: copy the tile from memory
    uint32_t bg_u8_splat;  // likewise four uint8 background pixels, valid when bg_splat != 0
};
constexpr int kTraceImages = 32768;

template <int R> struct Geo {
    static constexpr int Q = R - 2;             // densified image size (110 / 222)
    static constexpr int NS = R / 4;            // float4 strips per grid row
    static constexpr int MW = (R + 31) / 32;    // words of a row / column occupancy mask
    static constexpr int NSEG = R / 112;        // 28-strip segments per row in the dense pass
    static constexpr int HI_ROWS = R * R / S;   // source rows whose horizontal interpolation fits G
    static constexpr bool kImgSmem = R == 112;
};

template <int R> struct Smem {
    float G[R * R];        // one depth slice of the grid / pooled slice / horizontally interpolated rows
    float IMG[Geo<R>::kImgSmem ? (R - 2) * R : 4];   // running max over depth of the smoothed slices
    uint2 cache[CAP];      // per point: (x | y << 8 | slice << 16, value bits)
    unsigned ext[6];       // order-preserving keys of max xyz / min xyz
    unsigned rowmask[D][Geo<R>::MW], colmask[D][Geo<R>::MW];   // occupied grid rows / columns per slice
    int ylo[D], yhi[D], xlo[D], xhi[D];
    int cnt[D];            // points per slice (stamp path: counting sort)
    int ulo, uhi, vlo, vhi;   // rows / columns of IMG any slice can touch
    int oy_lo, oy_hi, g_lo, g_hi;   // output rows / 8-pixel column groups that can differ from background
    float red[NW];
    float mx;
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

// order-preserving float <-> unsigned keys (shared-memory atomicMax / atomicMin on floats of any sign)
__device__ __forceinline__ unsigned f2key(float f)
{
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(unsigned k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
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
template <int R>
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
    // `/ (1 + depth_bias)`: true division (torch-CPU, the oracle) or multiplication by the rounded
    // reciprocal (what torch-CUDA does for a division by a Python scalar), VgConfig.div_mode
    fz = P.div_mode == VG_DIV_RECIPROCAL ? __fmul_rn(fz, n.rc_opb)
                                         : div_rn(fz, P.one_plus_bias, n.rc_opb, n.slow);
    fz = __fmul_rn(fz, (float)(D - 2));
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
__device__ __forceinline__ void lin_idx(int Qn, int dst, int &i0, float &l0, float &l1)
{
    const float scale = __fdiv_rn((float)(Qn - 1), (float)(S - 1));
    const float src = __fmul_rn(scale, (float)dst);
    int a = (int)floorf(src);
    a = min(a, Qn - 1);
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

__global__ void projection_tables_kernel(ProjTables *t, int Qn)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid == 0) {
        unsigned nsm;
        asm volatile("mov.u32 %0, %%nsmid;" : "=r"(nsm));
        t->nsmid = (int)nsm;
    }
    if (tid < S) {
        int i0; float l0, l1;
        lin_idx(Qn, tid, i0, l0, l1);
        t->i0[tid] = i0; t->l0[tid] = l0; t->l1[tid] = l1;
    }
    // the image of an all-background (1.0) neighbourhood: HI = fma(1, lw0, 1*lw1), then the
    // vertical pass -- 254 or 255 depending on how (w0 + w1) rounds, exactly like torch-CPU
    for (int px = tid; px < S * S; px += gridDim.x * blockDim.x) {
        const int oy = px / S, ox = px - oy * S;
        int x0, y0; float lw0, lw1, lh0, lh1;
        lin_idx(Qn, ox, x0, lw0, lw1);
        lin_idx(Qn, oy, y0, lh0, lh1);
        const float c = __fmaf_rn(1.0f, lw0, __fmul_rn(1.0f, lw1));
        const float v = quant255(__fmaf_rn(c, lh0, __fmul_rn(c, lh1)));
        reinterpret_cast<uint8_t *>(t->bg_u8)[px] = (uint8_t)v;
        const int patch = (oy >> 4) * 14 + (ox >> 4), inner = (oy & 15) * 16 + (ox & 15);
        reinterpret_cast<op_t *>(t->bg_tile)[patch * 256 + inner] = to_op(v);
    }
}

// running depth-max image: shared memory (R = 112) or the per-SM global scratch (R = 224; .cg so
// that what another thread of the CTA wrote before the last barrier is what this load returns)
template <bool SMEM> __device__ __forceinline__ float4 img_ld4(const float *p)
{
    if (SMEM) return *reinterpret_cast<const float4 *>(p);
    return __ldcg(reinterpret_cast<const float4 *>(p));
}
template <bool SMEM> __device__ __forceinline__ void img_st4(float *p, float4 v)
{
    if (SMEM) *reinterpret_cast<float4 *>(p) = v;
    else __stcg(reinterpret_cast<float4 *>(p), v);
}
template <bool SMEM> __device__ __forceinline__ float img_ld(const float *p)
{
    if (SMEM) return *p;
    return __ldcg(p);
}

// a / b for 0 <= a < 2^15, 1 <= b <= 512 without the ~25-instruction integer division:
// floor((a + 0.5) / b) never crosses an integer, and the approximate reciprocal is good to 2 ulp
__device__ __forceinline__ int small_div(int a, int b)
{
    return __float2int_rz(__fdividef((float)a + 0.5f, (float)b));
}

// f(row, strip) over rows [r0, r1] x strips [s0, s1]: a warp per row, a lane per strip.  R = 112 has
// 28 strips, so a lane sees at most one strip of a row and the inner loop disappears.
template <int R, class F>
__device__ __forceinline__ void for_rows_strips(int r0, int r1, int s0, int s1, int warp, int lane, F f)
{
    if (R == 112) {
        const int s = s0 + lane;
        if (s <= s1)
            for (int r = r0 + warp; r <= r1; r += NW) f(r, s);
    } else {
        for (int r = r0 + warp; r <= r1; r += NW)
            for (int s = s0 + lane; s <= s1; s += 32) f(r, s);
    }
}

// 3x3 Gaussian (zero padding) of three pooled rows as a row-major FMA chain -- the summation order
// the oracle uses -- for the four pixels of one strip; t*[0] / t*[5] are the neighbouring columns
__device__ __forceinline__ void gauss4(const float (&w)[9], const float (&ta)[6], const float (&tb)[6],
                                       const float (&tc)[6], float (&o)[4])
{
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
}

template <int R>
__device__ __forceinline__ void project_image(const ProjParams &P, const int b, Smem<R> &sm)
{
    using GE = Geo<R>;
    constexpr int Q = GE::Q, NS = GE::NS, MW = GE::MW;
    constexpr bool ISM = GE::kImgSmem;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = b / P.V, v = b - c * P.V;
    const int beg = P.offsets[c];
    const int n = P.offsets[c + 1] - beg;
    const float *__restrict__ pts = P.points + 3 * (size_t)beg;
    const float *rm = P.rot + 9 * v;
    const bool fused = P.rotate_mode == VG_ROTATE_FUSED ||
                       (P.rotate_mode == VG_ROTATE_TORCH_CPU && 9 * (long long)n >= 400);
    float *IMG = sm.IMG;
    if (!ISM) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        IMG = P.img_scratch + (size_t)smid * Q * R;      // one resident CTA per SM at R = 224
    }

    // ---- phase 1: per-axis min / max of the rotated points --------------------------------------
    {
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
        if (tid < 6) sm.ext[tid] = tid < 3 ? 0u : 0xffffffffu;
        if (tid < D * MW) { (&sm.rowmask[0][0])[tid] = 0u; (&sm.colmask[0][0])[tid] = 0u; }
        if (tid < D) sm.cnt[tid] = 0;
        if (tid == 0) {
            sm.mask = 0u; sm.degenerate = 0;
            sm.ulo = Q; sm.uhi = -1; sm.vlo = Q; sm.vhi = -1;
            sm.oy_lo = S; sm.oy_hi = -1; sm.g_lo = NG; sm.g_hi = -1;
        }
        const unsigned k0 = __reduce_max_sync(0xffffffffu, f2key(mx0));
        const unsigned k1 = __reduce_max_sync(0xffffffffu, f2key(mx1));
        const unsigned k2 = __reduce_max_sync(0xffffffffu, f2key(mx2));
        const unsigned k3 = __reduce_min_sync(0xffffffffu, f2key(mn0));
        const unsigned k4 = __reduce_min_sync(0xffffffffu, f2key(mn1));
        const unsigned k5 = __reduce_min_sync(0xffffffffu, f2key(mn2));
        const bool all_finite = __all_sync(0xffffffffu, finite);
        __syncthreads();
        if (lane == 0) {
            atomicMax(&sm.ext[0], k0); atomicMax(&sm.ext[1], k1); atomicMax(&sm.ext[2], k2);
            atomicMin(&sm.ext[3], k3); atomicMin(&sm.ext[4], k4); atomicMin(&sm.ext[5], k5);
            if (!all_finite) sm.degenerate = 1;
        }
        __syncthreads();
    }
    Quant qn;
    {
        float a[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) a[k] = key2f(sm.ext[k]);
        qn.cx = __fmul_rn(__fadd_rn(a[0], a[3]), 0.5f);
        qn.cy = __fmul_rn(__fadd_rn(a[1], a[4]), 0.5f);
        qn.cz = __fmul_rn(__fadd_rn(a[2], a[5]), 0.5f);
        qn.pr = fmaxf(fmaxf(__fsub_rn(a[0], a[3]), __fsub_rn(a[1], a[4])), __fsub_rn(a[2], a[5]));
    }
    op_t *tile = P.tiles ? P.tiles + (size_t)b * VG_TILE_ELEMS : nullptr;
    uint8_t *u8 = nullptr;
    if (P.u8) {
        if (!P.u8_first_only) u8 = P.u8 + (size_t)b * S * S;
        else if (v == 0) u8 = P.u8 + (size_t)c * S * S;
    }
    if (sm.degenerate || n <= 0 || !(qn.pr > 0.0f) || !isfinite(qn.pr)) {
        // defined behaviour where the reference yields NaN: zero tile + status
        if (tile) {
            uint4 *t = reinterpret_cast<uint4 *>(tile);
            for (int i = tid; i < VG_TILE_ELEMS / 8; i += NT) t[i] = make_uint4(0, 0, 0, 0);
        }
        if (u8) {
            uint4 *t = reinterpret_cast<uint4 *>(u8);
            for (int i = tid; i < S * S / 16; i += NT) t[i] = make_uint4(0, 0, 0, 0);
        }
        if (P.status && v == 0 && tid == 0) P.status[c] = VG_EDEGENERATE;
        return;
    }
    if (P.status && v == 0 && tid == 0) P.status[c] = VG_OK;
    qn.rc_pr = __frcp_rn(qn.pr);
    qn.rc_opb = __frcp_rn(P.one_plus_bias);
    qn.slow = !(qn.pr > 1e-18f && qn.pr < 1e18f);

    // ---- phase 2: quantise once; occupied slices, rows and columns; cache (x, y, slice, value) ----
    // stamp path: every point stays in registers until the per-slice counts are known and is then
    // written to its slot of the slice-sorted cache.  dense path: the first CAP points go to the
    // cache in point order, the rest to a private pool in global memory (L2 resident; one per
    // resident CTA, claimed per SM), so no point is rotated and quantised more than once per view.
    const bool stamp = n <= STAMP_N && !P.dbg_grid;
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
        uint2 held[2];
        int hz[2] = {-1, -1}, hr[2] = {0, 0};
        auto one = [&](int i, int k) {
            float qx, qy, qz, val; int X, Y, zi;
            rotate_point(pts + 3 * i, rm, fused, qx, qy, qz);
            quantise<R>(qx, qy, qz, qn, P, X, Y, zi, val);
            m |= 1u << zi;
            atomicOr(&sm.rowmask[zi][Y >> 5], 1u << (Y & 31));
            atomicOr(&sm.colmask[zi][X >> 5], 1u << (X & 31));
            const uint2 e = make_uint2((unsigned)X | ((unsigned)Y << 8) | ((unsigned)zi << 16),
                                       __float_as_uint(val));
            if (stamp) {
                held[k] = e;
                hz[k] = zi;
                hr[k] = atomicAdd(&sm.cnt[zi], 1);
            } else if (i < CAP) {
                sm.cache[i] = e;
            } else if (i - CAP < npool) {
                __stcg(pool + (i - CAP), e);
            }
        };
        if (stamp) {
#pragma unroll
            for (int k = 0; k < 2; ++k)
                if (tid + k * NT < n) one(tid + k * NT, k);
        } else {
            for (int i = tid; i < n; i += NT) one(i, 0);
        }
        m = __reduce_or_sync(0xffffffffu, m);
        if (lane == 0 && m) atomicOr(&sm.mask, m);
        __syncthreads();
        if (stamp) {
#pragma unroll
            for (int k = 0; k < 2; ++k)
                if (hz[k] >= 0) {
                    int base = 0;
                    for (int d = 0; d < hz[k]; ++d) base += sm.cnt[d];
                    sm.cache[base + hr[k]] = held[k];
                }
        }
    }
    if (tid < 2 * D) {     // occupied row / column range per slice, and their union in IMG terms
        const int d = tid >> 1;
        const bool rows = (tid & 1) == 0;
        int lo = R, hi = -1;
#pragma unroll
        for (int w = 0; w < MW; ++w) {
            const unsigned bits = rows ? sm.rowmask[d][w] : sm.colmask[d][w];
            if (bits) {
                lo = min(lo, 32 * w + __ffs(bits) - 1);
                hi = max(hi, 32 * w + 31 - __clz(bits));
            }
        }
        if (rows) { sm.ylo[d] = lo; sm.yhi[d] = hi; } else { sm.xlo[d] = lo; sm.xhi[d] = hi; }
        if (hi >= 0) {
            atomicMin(rows ? &sm.ulo : &sm.vlo, max(lo - 4, 0));
            atomicMax(rows ? &sm.uhi : &sm.vhi, min(hi + 2, Q - 1));
        }
    }
    __syncthreads();
    const unsigned mask = sm.mask;
    const int ulo = sm.ulo, uhi = sm.uhi, vlo = sm.vlo, vhi = sm.vhi;   // IMG cells any slice writes
    const int nlo = max(ulo - 1, 0), nhi = min(uhi + 1, Q - 1);          // rows the emit may read
    // strips the emit may read: an 8-pixel output group is computed as soon as one of its pixels has a
    // source column in [vlo, vhi], and its other pixels reach up to kColSpan source columns further
    constexpr int kColSpan = (7 * (Q - 1)) / (S - 1) + 2;
    const int ns_lo = max(vlo - kColSpan, 0) >> 2, ns_hi = min(vhi + kColSpan, Q - 1) >> 2;
    const ProjTables *__restrict__ tab = P.tab;
    // output rows / 8-pixel column groups with at least one source pixel inside the touched region
    if (tid < S) {
        int y0; float t0, t1;
        lin_idx(Q, tid, y0, t0, t1);
        const int y1 = y0 + (y0 < Q - 1 ? 1 : 0);
        if (!(y1 < ulo || y0 > uhi)) { atomicMin(&sm.oy_lo, tid); atomicMax(&sm.oy_hi, tid); }
    } else if (tid < S + NG) {
        const int g = tid - S;
        const int xa = __ldg(&tab->i0[8 * g]);
        int xb = __ldg(&tab->i0[8 * g + 7]);
        xb += xb < Q - 1 ? 1 : 0;
        if (!(xb < vlo || xa > vhi)) { atomicMin(&sm.g_lo, g); atomicMax(&sm.g_hi, g); }
    }
    // IMG starts at 0 == max over the empty slices (their smoothed image is identically 0); the
    // dense pass writes whole rows, the stamp pass only inside [ns_lo, ns_hi]
    for_rows_strips<R>(nlo, nhi, stamp ? ns_lo : 0, stamp ? ns_hi : NS - 1, warp, lane, [&](int r, int s) {
        img_st4<ISM>(IMG + r * R + 4 * s, make_float4(0.f, 0.f, 0.f, 0.f));
    });
    if (!stamp)       // dense path: the one clear of the grid buffer
        for (int i = tid; i < R * R / 4; i += NT)
            reinterpret_cast<float4 *>(sm.G)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    const int oy_lo = sm.oy_lo, oy_hi = sm.oy_hi, g_lo = sm.g_lo, g_hi = sm.g_hi;

    // ---- background: every output group outside the active rows x groups is a copy of the
    // precomputed background tile; issued now so that the stores drain behind the stencil work ----
    // The tile is [196 patches][32 x 16 bytes]: a warp copies one whole patch (512 contiguous bytes) per
    // step, lane = (row inside the patch) * 2 + (half of the 16-pixel row); patches advance by 16 = one
    // patch row (14) + 2.
    {
        // lane-constant part of the activity test: oy = 16 py + ky in [oy_lo, oy_hi], g = 2 px + half in
        // [g_lo, g_hi]  <=>  py in [py_lo, py_hi] and px in [px_lo, px_hi] for this lane
        const int ky = lane >> 1, half = lane & 1;
        const int py_lo = (oy_lo - ky + 15) >> 4, py_hi = (oy_hi - ky) >> 4;     // arithmetic shifts: floor
        const int px_lo = (g_lo - half + 1) >> 1, px_hi = (g_hi - half) >> 1;
        const bool splat = P.bg_splat != 0u;     // the background is one value (R = 112, 224): nothing to load
        if (tile) {
            const uint4 *__restrict__ src = tab->bg_tile + lane;
            uint4 *dst = reinterpret_cast<uint4 *>(tile) + lane;
            const uint4 bgv = make_uint4(P.bg_splat, P.bg_splat, P.bg_splat, P.bg_splat);
            int py = warp >= 14 ? 1 : 0, px = warp >= 14 ? warp - 14 : warp;
#pragma unroll 13
            for (int p = warp; p < 196; p += NW) {
                if (py < py_lo || py > py_hi || px < px_lo || px > px_hi)
                    dst[p * 32] = splat ? bgv : __ldg(src + p * 32);
                px += 2; py += 1;
                if (px >= 14) { px -= 14; py += 1; }
            }
        }
        if (u8) {
            int py = warp >= 14 ? 1 : 0, px = warp >= 14 ? warp - 14 : warp;
            for (int p = warp; p < 196; p += NW) {
                const int oy = py * 16 + ky, g = px * 2 + half;
                if (py < py_lo || py > py_hi || px < px_lo || px > px_hi)
                    *reinterpret_cast<uint2 *>(u8 + oy * S + 8 * g) =
                        splat ? make_uint2(P.bg_u8_splat, P.bg_u8_splat) : __ldg(&tab->bg_u8[(oy * S + 8 * g) >> 3]);
                px += 2; py += 1;
                if (px >= 14) { px -= 14; py += 1; }
            }
        }
    }

    float w[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) w[k] = P.gauss[k];

    // ---- phase 3: per occupied slice: scatter-max, 5x5 max-pool, 3x3 Gaussian, depth max --------
    if (stamp) {
        int base = 0;
        for (int d = 0; d < D; ++d) {
            if (!((mask >> d) & 1u)) continue;
            const int cnt = sm.cnt[d];
            const int ylo = sm.ylo[d], yhi = sm.yhi[d], xlo = sm.xlo[d], xhi = sm.xhi[d];
            const int gy0 = max(ylo - 4, 0), gy1 = min(yhi + 2, Q - 1);       // smoothed rows
            const int gs_lo = max(xlo - 4, 0) >> 2, gs_hi = min(xhi + 2, Q - 1) >> 2;   // and strips
            const int py0 = max(ylo - 3, 0), py1 = min(yhi + 1, Q - 1);       // rows a stamp reaches
            const int ns = gs_hi - gs_lo + 1;
            // clear exactly the pooled cells this slice's Gaussian reads
            for_rows_strips<R>(py0, py1, gs_lo, gs_hi, warp, lane, [&](int r, int s) {
                reinterpret_cast<float4 *>(sm.G + r * R)[s] = make_float4(0.f, 0.f, 0.f, 0.f);
            });
            __syncthreads();
            // stamp: one item = one row of one point's 5x5 footprint.  Pooled pixel (q, x) covers raw
            // cells (q-1..q+3, x-1..x+3), so a point at (Y, X) reaches q in [Y-3, Y+1], x in [X-3, X+1].
            // All values are positive floats: integer order == float order.
            for (int item = tid; item < 5 * cnt; item += NT) {
                const int p = item / 5, dy = item - 5 * p;
                const uint2 e = sm.cache[base + p];
                const int X = (int)(e.x & 255u), Y = (int)((e.x >> 8) & 255u);
                const int q = Y - 3 + dy;
                if (q < 0 || q > Q - 1) continue;
                int *row = reinterpret_cast<int *>(sm.G) + q * R;
#pragma unroll
                for (int dx = 0; dx < 5; ++dx) {
                    const int x = X - 3 + dx;
                    if (x >= 0 && x <= Q - 1) atomicMax(row + x, (int)e.y);
                }
            }
            __syncthreads();
            // Gaussian + depth max over the bounding box: a row of ns strips is split into nseg
            // segments of `own` strips plus one halo lane on each side (neighbours come by shuffle),
            // rpw segment-rows share a warp.
            {
                const int nseg = ns > 30 ? 2 : 1;                    // R = 112: always 1 (ns <= 28)
                const int own = nseg == 1 ? ns : (ns + 1) >> 1;
                const int lw = own + 2;
                const int rpw = small_div(32, lw);
                const int rg = small_div(lane, lw), j = lane - rg * lw;
                const int tasks = (gy1 - gy0 + 1) * nseg;
                for (int t0 = warp * rpw; t0 < tasks; t0 += NW * rpw) {
                    const int t = t0 + rg;
                    const bool valid = rg < rpw && t < tasks;
                    const int rrow = nseg == 1 ? t : t >> 1, seg = nseg == 1 ? 0 : t & 1;
                    const int y = gy0 + rrow;
                    const int s = gs_lo + seg * own - 1 + j;
                    const bool inreg = valid && s >= gs_lo && s <= gs_hi;
                    const bool owner = inreg && j >= 1 && j <= own;
                    float ta[6], tb[6], tc[6];
                    auto load = [&](int rr, float (&tt)[6]) {
                        float4 cc = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (inreg && rr >= py0 && rr <= py1)
                            cc = reinterpret_cast<const float4 *>(sm.G + rr * R)[s];
                        tt[0] = __shfl_up_sync(0xffffffffu, cc.w, 1);
                        tt[5] = __shfl_down_sync(0xffffffffu, cc.x, 1);
                        tt[1] = cc.x; tt[2] = cc.y; tt[3] = cc.z; tt[4] = cc.w;
                    };
                    load(y - 1, ta);
                    load(y, tb);
                    load(y + 1, tc);
                    float o[4];
                    gauss4(w, ta, tb, tc, o);
                    if (s == NS - 1) { o[2] = 0.0f; o[3] = 0.0f; }      // columns Q, Q+1 are padding
                    if (owner) {
                        float *dst = IMG + y * R + 4 * s;
                        img_st4<ISM>(dst, max4(img_ld4<ISM>(dst), make_float4(o[0], o[1], o[2], o[3])));
                    }
                }
            }
            __syncthreads();
            base += cnt;
        }
    } else {
        // Dense path.  The grid buffer is cleared ONCE.  Slices are visited in increasing depth and a
        // point of slice d carries a value in (d-1, d], so whatever lower slices left behind is <= d-1:
        // scatter-max simply overwrites it, and after pooling (max commutes with the monotone cut)
        // everything <= d-1 is cut to the 0 an empty cell holds.  Only slices whose values can tie with
        // a lower one (0/1: both 1.0, 7: 6.0 like the top of slice 6 -- both only reachable through
        // rounding) clear again.  Per slice that leaves two block barriers: scatter | fused stencil
        // pass.  In the fused pass a warp owns a chunk of smoothed rows of one 28-strip segment and
        // streams the raw rows it needs through registers: horizontal 5-max (shuffles), vertical 5-max
        // over a ring of 5 rows, cut, 3x3 Gaussian over a ring of 3 pooled rows, running depth max into
        // IMG.  Rows outside [ylo, yhi] hold no point of the slice and are not even loaded.
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
                auto put = [&](const uint2 e) {
                    if ((int)(e.x >> 16) == d)
                        atomicMax(reinterpret_cast<int *>(sm.G) + (int)((e.x >> 8) & 255u) * R + (int)(e.x & 255u),
                                  (int)e.y);
                };
                const int ncache = min(n, CAP);
                for (int i = tid; i < ncache; i += NT) put(sm.cache[i]);
                // pooled points: 4 independent L2 loads in flight per thread
                for (int i0 = tid; i0 < npool; i0 += 4 * NT) {
                    uint2 e[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        e[k] = i0 + k * NT < npool ? __ldcg(pool + i0 + k * NT) : make_uint2(0xffffffffu, 0u);
#pragma unroll
                    for (int k = 0; k < 4; ++k) put(e[k]);
                }
                // beyond the pool (N > CAP + POOL), or every pool of this SM taken (cannot happen at
                // two resident CTAs per SM): rotate and quantise again
                for (int i = CAP + npool + tid; i < n; i += NT) {
                    float qx, qy, qz, val; int X, Y, zi;
                    rotate_point(pts + 3 * i, rm, fused, qx, qy, qz);
                    quantise<R>(qx, qy, qz, qn, P, X, Y, zi, val);
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
                constexpr int NSEG = GE::NSEG;
                const int glo = max(ylo - 4, 0), ghi = min(yhi + 2, Q - 1);   // smoothed rows of this slice
                const int rows = ghi - glo + 1;
                const int ch = min(max((rows * NSEG + NW - 1) / NW, 2), CH);  // rows per task, 2..7
                const int nchunks = (rows + ch - 1) / ch;
                for (int task = warp; task < nchunks * NSEG; task += NW) {
                    const int chunk = task / NSEG, seg = task - chunk * NSEG;
                    const int y0 = glo + chunk * ch;
                    const int y1 = min(y0 + ch - 1, ghi);
                    // segment 0: strips 0..31, lanes 0..27 own; segment 1 (R = 224): strips 24..55,
                    // lanes 4..31 own (the other lanes only feed their neighbours)
                    const int strip = seg * 24 + lane;
                    const bool owner = strip < NS && (seg == 0 ? lane < 28 : lane >= 4);
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
                        // H(r, x) = max G(r, x-1 .. x+3) for x in [0, Q); columns Q, Q+1 are padding
                        float4 h = z4;
                        if (r >= ylo && r <= yhi) {            // warp-uniform: other rows hold no point of d
                            float4 a = z4;
                            if (strip < NS) a = reinterpret_cast<const float4 *>(sm.G + r * R)[strip];
                            const float left = __shfl_up_sync(0xffffffffu, a.w, 1);
                            const float rx = __shfl_down_sync(0xffffffffu, a.x, 1);
                            const float ry = __shfl_down_sync(0xffffffffu, a.y, 1);
                            const float rz = __shfl_down_sync(0xffffffffu, a.z, 1);
                            const float l = strip == 0 ? 0.0f : left;
                            const float m3 = max3f(a.z, a.w, rx);
                            h.x = max3f(max3f(l, a.x, a.y), a.z, a.w);
                            h.y = max3f(m3, a.x, a.y);
                            h.z = max3f(m3, a.y, ry);
                            h.w = max3f(m3, ry, rz);
                            if (strip == NS - 1) { h.z = 0.0f; h.w = 0.0f; }
                        }
                        h0 = h1; h1 = h2; h2 = h3; h3 = h4; h4 = h;
                        if (i < 4) continue;
                        // P(q, x) = max H(q-1 .. q+3, x), q = r - 3; rows Q, Q+1 are padding; cut the
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
                        if (strip == 0) lc = 0.0f;
                        if (i < 6) continue;
                        // 3x3 Gaussian (zero padding) of pooled rows y-1, y, y+1 (y = r - 4), then the
                        // running max over depth
                        const int y = r - 4;
                        const float ta[6] = {la, pa.x, pa.y, pa.z, pa.w, ra};
                        const float tb[6] = {lb, pb.x, pb.y, pb.z, pb.w, rb};
                        const float tc[6] = {lc, pc.x, pc.y, pc.z, pc.w, rc};
                        float o[4];
                        gauss4(w, ta, tb, tc, o);
                        if (strip == NS - 1) { o[2] = 0.0f; o[3] = 0.0f; }
                        if (owner) {
                            float *dst = IMG + y * R + 4 * strip;
                            img_st4<ISM>(dst, max4(img_ld4<ISM>(dst), make_float4(o[0], o[1], o[2], o[3])));
                        }
                    }
                }
            }
            __syncthreads();
        }
    }

    if (big) {   // hand the pool back (every read of it is behind the last slice's barriers)
        if (tid == 0 && sm.spill_slot >= 0) {
            __threadfence();
            atomicExch(&P.spill_flags[sm.spill_slot], 0);
        }
    }

    // ---- phase 4: img / max(img), 1 - x  (cells the emit reads; everything else is background) ----
    {
        const int us_lo = vlo >> 2, us_hi = vhi >> 2;
        float m = 0.0f;
        for_rows_strips<R>(ulo, uhi, us_lo, us_hi, warp, lane, [&](int r, int s) {
            const float4 t = img_ld4<ISM>(IMG + r * R + 4 * s);
            m = fmaxf(fmaxf(m, fmaxf(t.x, t.y)), fmaxf(t.z, t.w));
        });
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
        for_rows_strips<R>(nlo, nhi, ns_lo, ns_hi, warp, lane, [&](int r, int s) {
            float4 t = img_ld4<ISM>(IMG + r * R + 4 * s);
            t.x = __fsub_rn(1.0f, div_rn(t.x, mx, rc_mx, false));
            t.y = __fsub_rn(1.0f, div_rn(t.y, mx, rc_mx, false));
            t.z = __fsub_rn(1.0f, div_rn(t.z, mx, rc_mx, false));
            t.w = __fsub_rn(1.0f, div_rn(t.w, mx, rc_mx, false));
            img_st4<ISM>(IMG + r * R + 4 * s, t);
        });
        __syncthreads();
        if (P.dbg_dens) {
            float *dd = P.dbg_dens + (size_t)b * Q * Q;
            for (int i = tid; i < Q * Q; i += NT) {
                const int y = i / Q, x = i - y * Q;
                const bool in = y >= nlo && y <= nhi && (x >> 2) >= ns_lo && (x >> 2) <= ns_hi;
                dd[i] = in ? img_ld<ISM>(IMG + y * R + x) : 1.0f;
            }
        }
    }

    // ---- phase 5: bilinear Q -> 224 (align_corners), floor(x*255), patch-major tiles ---------------
    // HI[y][ox] = fma(IMG[y][x0], lw0, IMG[y][x1]*lw1) for up to HI_ROWS source rows at a time (held
    // in the grid buffer), then out = fma(HI[y0], lh0, HI[y1]*lh1) -- the exact contraction pattern of
    // torch-CPU's separable interpolation on an FMA host.  Only the active rows x column groups are
    // computed (the rest was copied from the background tile above); threads map to a fixed output
    // column (horizontal pass) / column group (vertical pass) and stride over the rows.
    if (oy_hi < oy_lo || g_hi < g_lo) return;          // cannot happen for a non-degenerate cluster
    {
        constexpr int HI_ROWS = GE::HI_ROWS;
        const int ng = g_hi - g_lo + 1, ncols = 8 * ng;
        const int h_nrl = small_div(NT, ncols), h_rl = small_div(tid, ncols);   // horizontal pass: row lanes
        const int ox = 8 * g_lo + (tid - h_rl * ncols);
        int cx0 = 0, cx1 = 0;
        float clw0 = 0.f, clw1 = 0.f;
        if (h_rl < h_nrl) {
            cx0 = __ldg(&tab->i0[ox]);
            clw0 = __ldg(&tab->l0[ox]);
            clw1 = __ldg(&tab->l1[ox]);
            cx1 = cx0 + (cx0 < Q - 1 ? 1 : 0);
        }
        const int v_nrl = small_div(NT, ng), v_rl = small_div(tid, ng);         // vertical pass: row lanes
        const int g = g_lo + (tid - v_rl * ng);
        const f32x2 k255 = pack2(255.0f, 255.0f), kmagic = pack2(8388608.0f, 8388608.0f);
        const float scale = __fdiv_rn((float)(Q - 1), (float)(S - 1));
        int oy_cur = oy_lo;
        while (oy_cur <= oy_hi) {
            int ybase; float t0, t1;
            lin_idx(Q, oy_cur, ybase, t0, t1);
            const int ylast = min(ybase + HI_ROWS - 1, nhi);
            int oy_end = oy_hi;
            if (ylast < nhi) {
                // last output row whose two source rows are both inside [ybase, ylast]
                int gss = min(max((int)((float)ylast / scale), oy_cur), oy_hi);
                auto y1_of = [&](int oy) { int a; float u0, u1; lin_idx(Q, oy, a, u0, u1); return a + (a < Q - 1 ? 1 : 0); };
                while (gss + 1 <= oy_hi && y1_of(gss + 1) <= ylast) ++gss;
                while (gss > oy_cur && y1_of(gss) > ylast) --gss;
                oy_end = gss;
            }
            if (h_rl < h_nrl)
                for (int y = ybase + h_rl; y <= ylast; y += h_nrl) {
                    const float *row = IMG + y * R;
                    sm.G[(y - ybase) * S + ox] =
                        __fmaf_rn(img_ld<ISM>(row + cx0), clw0, __fmul_rn(img_ld<ISM>(row + cx1), clw1));
                }
            __syncthreads();
            if (v_rl < v_nrl)
                for (int oy = oy_cur + v_rl; oy <= oy_end; oy += v_nrl) {
                    int y0; float lh0, lh1;
                    lin_idx(Q, oy, y0, lh0, lh1);     // same arithmetic as the column table (bit-identical)
                    const int y1 = y0 + (y0 < Q - 1 ? 1 : 0);
                    const int patch = (oy >> 4) * 14 + (g >> 1);
                    const int inner = (oy & 15) * 16 + (g & 1) * 8;
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
#ifndef VG_OPERAND_BF16
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
            oy_cur = oy_end + 1;
            if (oy_cur <= oy_hi) __syncthreads();      // the next pass overwrites the HI rows
        }
    }
}

// one image per CTA, or -- list mode -- the images the fast kernel handed over, a fixed grid striding
// over the list
template <int R, bool LIST>
__global__ void __launch_bounds__(NT, R == 112 ? 2 : 1) projection_kernel(const ProjParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem<R> &sm = *reinterpret_cast<Smem<R> *>(smem_raw);
    if (!LIST) {
        project_image<R>(P, (int)blockIdx.x, sm);
        return;
    }
    // images are drawn from the list one at a time (they differ a lot in cost): defer[1] is the cursor
    const int count = P.defer[0];
    for (;;) {
        if (threadIdx.x == 0) sm.spill_slot = atomicAdd(&P.defer[1], 1);
        __syncthreads();
        const int i = sm.spill_slot;
        __syncthreads();       // everybody has the index before project_image reuses the field
        if (i >= count) break;
        project_image<R>(P, P.defer[3 + i], sm);
        __syncthreads();       // the next image re-initialises the shared state
    }
}

// =================================================================================================
// Fast path: R = 112, clusters of up to FAST_N points (every cluster of a Waymo-shaped frame), images
// whose touched region is at most F_MAXR rows x F_MAXS float4 strips (always, at the reference's
// obj_ratio = 0.8: X, Y in [12, 101]).  Same arithmetic as project_image's stamp path -- the results
// are bit-identical -- organised for instruction count:
//   * 256-thread CTAs, 74 KB of shared memory -> three CTAs per SM: half the per-warp replicated
//     overhead of the 512-thread kernel, and three independent barrier domains per SM.
//   * the rotated points wait in the (still unused) slice buffer between the min / max pass and the
//     quantisation: rolled point loops, small code.
//   * ONE bounding-box-limited buffer B (origin = the union of all slices' touched regions, fixed
//     pitch, zero halo) holds the pooled slice; the running depth-max image lives in REGISTERS: a
//     thread owns up to KQ (row pair, strip) items of the union for the whole image, so a slice costs
//     no image load / store and no zero-initialisation pass, and a row pair shares two of its four
//     pooled rows.  Neighbour columns arrive by shuffle (row starts / warp edges: one predicated load
//     from the zero halo / the neighbouring warp's strip).
//   * after the last slice the normalised image is written back into B (margins = 1.0) and the
//     bilinear emit WALKS: a thread owns four adjacent output columns and a run of source rows, keeps the
//     horizontally interpolated source rows y0 / y0 + 1 in registers and advances them as the run
//     moves down -- no intermediate buffer, no barrier, every horizontal interpolation done ~once.
//     Its tables are laid out for the loop: the column weights as the packed pairs FFMA2 consumes (one
//     8-byte load, nothing to re-pack per row) and one 16-byte record per output row (both weights, the
//     next row's source row, the row's byte offset in the tile); when only tiles are wanted the row loop
//     carries no pointer tests: 21 instead of 34 instructions per emitted row of four pixels.  The
//     horizontal pass loads the run of 4 (5) source floats its four columns draw from ONCE and picks the
//     operands with per-thread byte-permute selectors: half the shared-memory wavefronts of eight scalar
//     loads whose lanes are two floats apart.
constexpr int FAST_N = 2048;
// largest touched region the path takes (rows x float4 strips) and its buffer.  R = 112: 96 x 24 (74 KB per
// CTA, three 256-thread CTAs per SM).  R = 224 (BASELINE configs[3]): 186 x 48 (X, Y in [23, 202]), 187 KB:
// one 512-thread CTA per SM with nine (row pair, strip) items per thread in its 128 registers.
template <int R> struct FastGeo {
    static constexpr int MAXR = R == 112 ? 96 : 186, MAXS = R == 112 ? 24 : 48;
    static constexpr int PITCH = MAXS + 5;     // float4 strips per buffer row: two halo strips per side, odd
    static constexpr int BROWS = MAXR + 3;     // halo row below / above + the second row of an odd last pair
};

template <int R> struct FastSmem {
    float4 B[FastGeo<R>::BROWS * FastGeo<R>::PITCH];
    uint2 cache[FAST_N];   // per point, sorted by depth slice: (x | y << 8, value bits)
    unsigned ext[6];
    unsigned rowmask[D][(R + 31) / 32], colmask[D][(R + 31) / 32];   // occupied grid rows / columns per slice
    int ylo[D], yhi[D], xlo[D], xhi[D], cnt[D], base[D];
    int ulo, uhi, vlo, vhi;
    float red[32];
    int degenerate;
    int next;              // next image of this CTA (persistent loop)
    uint4 bgsrc[448];      // one patch row (14 patches x 512 B) of background pixels: source of the bulk stores
    union {
        float2 lw[S];      // TAB = 0: bilinear weights (l0, l1) of output row / column i
        struct {           // TAB = 1: the column weights as the packed pairs the emit multiplies with:
            float2 cw0[S / 2], cw1[S / 2];     // cw0[k] = (l0[2k], l0[2k + 1]), cw1[k] = (l1[2k], l1[2k + 1])
        };
    };
    float4 rowrec[S];      // TAB = 1, per output row: (l0, l1, first source row of the NEXT output row, byte
                           // offset of the row inside a patch-major tile) -- one 16-byte load per emitted row
    unsigned char i0[S];   // first source row / column of output row / column i (copy of ProjTables, per CTA)
};
// three CTAs per SM at R = 112: 228 KB of shared memory per SM, 1 KB reserved per CTA
static_assert(sizeof(FastSmem<112>) <= (233472 / 3 - 1024), "FastSmem<112> must leave room for three CTAs per SM");
static_assert(sizeof(FastSmem<224>) <= 232448, "FastSmem<224> exceeds the shared memory of an SM");

// shared -> global bulk copy (TMA, no tensor map): the copy engine reads shared memory and writes L2
// without passing through the load/store unit
__device__ __forceinline__ void bulk_store(void *gdst, unsigned ssrc, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// phase timeline of the fast kernel: compiled in only with -DVG_PROJ_TRACE (python -m vilgod_b200.build
// --trace -> libvilgod_b200_trace.so); the counters would otherwise cost registers in the slice loop
#ifdef VG_PROJ_TRACE
#define VG_TR(k)                                                                   \
    do {                                                                           \
        if (P.trace && tid == 0 && b < kTraceImages) P.trace[b * 12 + (k)] = clock64(); \
    } while (0)
#define VG_TR_ON(x) x
#else
#define VG_TR(k) do { } while (0)
#define VG_TR_ON(x)
#endif

// one image (cluster b / V, view b % V); the bilinear table and the constant background row in `sm` were
// filled by the kernel before its image loop
template <int R, int NTF, int CW, int STAMP, int TAB>
__device__ __forceinline__ void fast_image(const ProjParams &P, const int b, FastSmem<R> &sm)
{
    constexpr int Q = R - 2, NS = R / 4, MW = (R + 31) / 32;
    constexpr int F_MAXR = FastGeo<R>::MAXR, F_MAXS = FastGeo<R>::MAXS, F_PITCH = FastGeo<R>::PITCH;
    constexpr int NWF = NTF / 32;
    constexpr int KQ = ((F_MAXR / 2) * F_MAXS + NTF - 1) / NTF;         // (row pair, strip) items per thread
    constexpr int PW = 4 * F_PITCH;                                     // buffer pitch in floats
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = b / P.V, v = b - c * P.V;
    // shared state first, and ONE early barrier (every warp reaches it at once) instead of one between
    // the reduction's initial values and its atomics
    if (tid < 6) sm.ext[tid] = tid < 3 ? 0u : 0xffffffffu;
    if (tid < D * MW) { (&sm.rowmask[0][0])[tid] = 0u; (&sm.colmask[0][0])[tid] = 0u; }
    if (tid < D) sm.cnt[tid] = 0;
    if (tid == 0) {
        sm.degenerate = 0;
        sm.ulo = Q; sm.uhi = -1; sm.vlo = Q; sm.vhi = -1;
    }
    const int beg = P.offsets[c];
    const int n = P.offsets[c + 1] - beg;
    __syncthreads();
    auto hand_over = [&]() {
        if (tid == 0) P.defer[3 + atomicAdd(P.defer, 1)] = b;
    };
    if (n > FAST_N) { hand_over(); return; }
    VG_TR(0);
    const float *__restrict__ pts = P.points + 3 * (size_t)beg;
    const float *rm = P.rot + 9 * v;
    const bool fused = P.rotate_mode == VG_ROTATE_FUSED ||
                       (P.rotate_mode == VG_ROTATE_TORCH_CPU && 9 * (long long)n >= 400);

    // ---- phase 1: rotate, per-axis min / max.  The rotated points wait in shared memory (the buffer B
    // is free until the first slice) so that this loop and the next one stay rolled: small code, no
    // register arrays ---------------------------------------------------------------------------------
    float *qxs = reinterpret_cast<float *>(sm.B), *qys = qxs + FAST_N, *qzs = qys + FAST_N;
    uint2 *tmp = reinterpret_cast<uint2 *>(qzs + FAST_N);   // unsorted (x | y << 8 | slice << 16 | rank << 19, value)
    static_assert(FAST_N * 20 <= sizeof(sm.B), "staging of the points must fit the slice buffer");
    {
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY;
        float mn0 = INFINITY, mn1 = INFINITY, mn2 = INFINITY;
        bool finite = true;
#pragma unroll 1
        for (int i = tid; i < n; i += NTF) {
            float qx, qy, qz;
            rotate_point(pts + 3 * i, rm, fused, qx, qy, qz);
            qxs[i] = qx; qys[i] = qy; qzs[i] = qz;
            finite = finite && isfinite(qx) && isfinite(qy) && isfinite(qz);
            mx0 = fmaxf(mx0, qx); mx1 = fmaxf(mx1, qy); mx2 = fmaxf(mx2, qz);
            mn0 = fminf(mn0, qx); mn1 = fminf(mn1, qy); mn2 = fminf(mn2, qz);
        }
        const unsigned k0 = __reduce_max_sync(0xffffffffu, f2key(mx0));
        const unsigned k1 = __reduce_max_sync(0xffffffffu, f2key(mx1));
        const unsigned k2 = __reduce_max_sync(0xffffffffu, f2key(mx2));
        const unsigned k3 = __reduce_min_sync(0xffffffffu, f2key(mn0));
        const unsigned k4 = __reduce_min_sync(0xffffffffu, f2key(mn1));
        const unsigned k5 = __reduce_min_sync(0xffffffffu, f2key(mn2));
        const bool all_finite = __all_sync(0xffffffffu, finite);
        if (lane == 0) {
            atomicMax(&sm.ext[0], k0); atomicMax(&sm.ext[1], k1); atomicMax(&sm.ext[2], k2);
            atomicMin(&sm.ext[3], k3); atomicMin(&sm.ext[4], k4); atomicMin(&sm.ext[5], k5);
            if (!all_finite) sm.degenerate = 1;
        }
        __syncthreads();
    }
    Quant qn;
    {
        float a[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) a[k] = key2f(sm.ext[k]);
        qn.cx = __fmul_rn(__fadd_rn(a[0], a[3]), 0.5f);
        qn.cy = __fmul_rn(__fadd_rn(a[1], a[4]), 0.5f);
        qn.cz = __fmul_rn(__fadd_rn(a[2], a[5]), 0.5f);
        qn.pr = fmaxf(fmaxf(__fsub_rn(a[0], a[3]), __fsub_rn(a[1], a[4])), __fsub_rn(a[2], a[5]));
    }
    op_t *tile = P.tiles ? P.tiles + (size_t)b * VG_TILE_ELEMS : nullptr;
    uint8_t *u8 = nullptr;
    if (P.u8) {
        if (!P.u8_first_only) u8 = P.u8 + (size_t)b * S * S;
        else if (v == 0) u8 = P.u8 + (size_t)c * S * S;
    }
    if (sm.degenerate || n <= 0 || !(qn.pr > 0.0f) || !isfinite(qn.pr)) {
        if (tile) {
            uint4 *t = reinterpret_cast<uint4 *>(tile);
            for (int i = tid; i < VG_TILE_ELEMS / 8; i += NTF) t[i] = make_uint4(0, 0, 0, 0);
        }
        if (u8) {
            uint4 *t = reinterpret_cast<uint4 *>(u8);
            for (int i = tid; i < S * S / 16; i += NTF) t[i] = make_uint4(0, 0, 0, 0);
        }
        if (P.status && v == 0 && tid == 0) P.status[c] = VG_EDEGENERATE;
        return;
    }
    qn.rc_pr = __frcp_rn(qn.pr);
    qn.rc_opb = __frcp_rn(P.one_plus_bias);
    qn.slow = !(qn.pr > 1e-18f && qn.pr < 1e18f);
    VG_TR(1);

    // ---- phase 2: quantise; per-slice counts and occupied rows / columns; counting sort by slice ----
#pragma unroll 1
    for (int i = tid; i < n; i += NTF) {
        float val; int X, Y, zi;
        quantise<R>(qxs[i], qys[i], qzs[i], qn, P, X, Y, zi, val);
        atomicOr(&sm.rowmask[zi][Y >> 5], 1u << (Y & 31));
        atomicOr(&sm.colmask[zi][X >> 5], 1u << (X & 31));
        const unsigned rank = (unsigned)atomicAdd(&sm.cnt[zi], 1);
        tmp[i] = make_uint2((unsigned)X | ((unsigned)Y << 8) | ((unsigned)zi << 16) | (rank << 19),
                            __float_as_uint(val));
    }
    __syncthreads();
    VG_TR(2);
    if (tid < 2 * D) {
        const int d = tid >> 1;
        const bool rows = (tid & 1) == 0;
        int lo = R, hi = -1;
#pragma unroll
        for (int w = 0; w < MW; ++w) {
            const unsigned bits = rows ? sm.rowmask[d][w] : sm.colmask[d][w];
            if (bits) {
                lo = min(lo, 32 * w + __ffs(bits) - 1);
                hi = max(hi, 32 * w + 31 - __clz(bits));
            }
        }
        if (rows) { sm.ylo[d] = lo; sm.yhi[d] = hi; } else { sm.xlo[d] = lo; sm.xhi[d] = hi; }
        if (hi >= 0) {
            atomicMin(rows ? &sm.ulo : &sm.vlo, max(lo - 4, 0));
            atomicMax(rows ? &sm.uhi : &sm.vhi, min(hi + 2, Q - 1));
        }
    } else if (tid >= 32 && tid < 32 + D) {
        int acc = 0;
        for (int d = 0; d < tid - 32; ++d) acc += sm.cnt[d];
        sm.base[tid - 32] = acc;
    }
    {
        // counting sort: every thread forms the slice offsets itself (eight 16-bit prefixes in two
        // registers), so the sorted cache is complete at the same barrier as the ranges
        unsigned long long plo = 0ull, phi = 0ull;
        unsigned acc = 0u;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            if (d < 4) plo |= (unsigned long long)acc << (16 * d);
            else phi |= (unsigned long long)acc << (16 * (d - 4));
            acc += (unsigned)sm.cnt[d];
        }
#pragma unroll 1
        for (int i = tid; i < n; i += NTF) {
            const uint2 e = tmp[i];
            const unsigned zi = (e.x >> 16) & 7u;
            const unsigned base = (unsigned)(((zi & 4u) ? phi : plo) >> (16 * (zi & 3u))) & 0xffffu;
            sm.cache[base + (e.x >> 19)] = make_uint2(e.x & 0xffffu, e.y);
        }
    }
    __syncthreads();
    VG_TR(3);
    const int ulo = sm.ulo, uhi = sm.uhi, vlo = sm.vlo, vhi = sm.vhi;   // image cells any slice writes
    const int nr = uhi - ulo + 1;
    const int us_lo = vlo >> 2, ns = (vhi >> 2) - us_lo + 1;
    const int npair = ((nr + 1) >> 1) * ns;
    if (nr > F_MAXR || ns > F_MAXS || npair > KQ * NTF) { hand_over(); return; }    // block-uniform
    if (P.status && v == 0 && tid == 0) P.status[c] = VG_OK;
    const ProjTables *__restrict__ tab = P.tab;
    // Output rows / 8-pixel column groups with a source pixel inside the touched region, in closed form:
    // the first source row of output row oy is i0[oy] = floor(oy (Q-1) / (S-1)) (projection_init checks
    // the table against this), the second one i0 + 1, so
    //   row oy is active     <=>  ulo - 1 <= i0[oy] <= uhi
    //   group g is active    <=>  i0[8 g + 7] >= vlo - 1  and  i0[8 g] <= vhi
    // and i0[o] >= t  <=>  o >= ceil(t (S-1) / (Q-1)).
    auto first_with = [](int t) { return t <= 0 ? 0 : (t * (S - 1) + (Q - 2)) / (Q - 1); };
    const int oy_lo = first_with(ulo - 1), oy_hi = min(first_with(uhi + 1) - 1, S - 1);
    const int g_lo = first_with(vlo - 1) >> 3, g_hi = min(min(first_with(vhi + 1) - 1, S - 1) >> 3, NG - 1);
    // buffer coordinates: grid row y -> y + orow, grid column x -> x + ocol (float index inside a row)
    const int orow = 1 - ulo, ocol = 4 * (2 - us_lo);

    // ---- background: every 16-byte piece of the tile outside the active rows x column groups holds the
    // background value.  R = 112 (one value everywhere): whole patch rows above / below the active patches
    // and the runs of patches left / right of them go out as bulk stores from a constant patch row in
    // shared memory (at most two per patch row, issued by one lane per warp: ~80 KB per image that never touch the
    // load/store unit); only the pieces of the border patches are written by the lanes.  Otherwise
    // (background tile with a pattern): a warp owns the patch columns px = warp, warp + NWF, a lane the
    // piece (row ky, half) of a patch, and copies from the precomputed tile --------------------------------
    {
        const int ky = lane >> 1, half = lane & 1;
        const bool splat = P.bg_splat != 0u;
        constexpr int NPX = (14 + NWF - 1) / NWF;
        if (tile && splat) {
            const int py_a = oy_lo >> 4, py_b = oy_hi >> 4, px_a = g_lo >> 1, px_b = g_hi >> 1;
            if (lane == 0) {                // one issuing thread per warp, patch rows py = warp, warp + NWF
                fence_async_smem();         // the constant row was written through the generic proxy
                const unsigned src = (unsigned)__cvta_generic_to_shared(sm.bgsrc);
                for (int py = warp; py < 14; py += NWF) {
                    char *rowp = reinterpret_cast<char *>(tile) + py * 7168;
                    if (py < py_a || py > py_b) {
                        bulk_store(rowp, src, 7168u);
                    } else {
                        if (px_a > 0) bulk_store(rowp, src, (unsigned)px_a * 512u);
                        if (px_b < 13) bulk_store(rowp + (px_b + 1) * 512, src, (unsigned)(13 - px_b) * 512u);
                    }
                }
                bulk_commit();
            }
            uint4 *dst = reinterpret_cast<uint4 *>(tile) + lane;
            const uint4 bgv = make_uint4(P.bg_splat, P.bg_splat, P.bg_splat, P.bg_splat);
            auto piece = [&](int py, int px) {
                const int oy = 16 * py + ky, gg = 2 * px + half;
                if (!(oy >= oy_lo && oy <= oy_hi && gg >= g_lo && gg <= g_hi)) dst[(py * 14 + px) * 32] = bgv;
            };
            for (int px = px_a + warp; px <= px_b; px += NWF) {          // top / bottom border patches
                piece(py_a, px);
                if (py_b != py_a) piece(py_b, px);
            }
            for (int py = py_a + 1 + warp; py < py_b; py += NWF) {       // left / right border patches
                piece(py, px_a);
                if (px_b != px_a) piece(py, px_b);
            }
        } else if (tile) {
            bool colact[NPX];
#pragma unroll
            for (int j = 0; j < NPX; ++j) {
                const int gg = 2 * (warp + j * NWF) + half;
                colact[j] = gg >= g_lo && gg <= g_hi;
            }
            const uint4 *__restrict__ src = tab->bg_tile + tid;
            uint4 *dst = reinterpret_cast<uint4 *>(tile) + tid;
#pragma unroll 7
            for (int py = 0; py < 14; ++py) {
                const int oy = 16 * py + ky;
                const bool rowact = oy >= oy_lo && oy <= oy_hi;
#pragma unroll
                for (int j = 0; j < NPX; ++j)
                    if (warp + j * NWF < 14 && !(rowact && colact[j]))
                        dst[py * 448 + j * NTF] = __ldg(src + py * 448 + j * NTF);
            }
        }
        if (u8) {
            for (int py = 0; py < 14; ++py) {
                const int oy = 16 * py + ky;
                const bool rowact = oy >= oy_lo && oy <= oy_hi;
#pragma unroll
                for (int j = 0; j < NPX; ++j) {
                    const int gg = 2 * (warp + j * NWF) + half;
                    if (warp + j * NWF < 14 && !(rowact && gg >= g_lo && gg <= g_hi))
                        *reinterpret_cast<uint2 *>(u8 + oy * S + 8 * gg) =
                            splat ? make_uint2(P.bg_u8_splat, P.bg_u8_splat)
                                  : __ldg(&tab->bg_u8[(oy * S + 8 * gg) >> 3]);
                }
            }
        }
    }

    VG_TR(4);
    // ---- phase 3: per occupied slice: stamp the 5x5 footprints, 3x3 Gaussian, depth max ------------
    // item k of a thread: rows (2 yp, 2 yp + 1) of the union x strip su; img0 / img1 = their running maxima
    float4 img0[KQ], img1[KQ];
    int boff[KQ];              // float index in B of the first row's strip; bit 29: take the left /
                               // bit 30: the right neighbour column from memory instead of by shuffle
#pragma unroll
    for (int k = 0; k < KQ; ++k) {
        const int item = tid + k * NTF;
        const int yp = small_div(item, ns), su = item - yp * ns;
        int o = ((2 * yp + 1) * F_PITCH + su + 2) * 4;
        if (lane == 0 || su == 0) o |= 1 << 29;
        if (lane == 31 || su == ns - 1) o |= 1 << 30;
        boff[k] = o;
        img0[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        img1[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float (&w)[9] = P.gauss;
    VG_TR(5);
    VG_TR_ON(long long tr_clear = 0; long long tr_stamp = 0; long long tr_gauss = 0; long long tr_t = 0;)
    unsigned occ = 0u;         // occupied depth slices
#pragma unroll
    for (int d = 0; d < D; ++d) occ |= sm.cnt[d] > 0 ? 1u << d : 0u;
    bool first = true;
    for (; occ; occ &= occ - 1) {
        const int d = __ffs(occ) - 1;
        const int cnt = sm.cnt[d];
        if (!first) __syncthreads();       // every Gaussian read of the previous slice is done
        first = false;
        const int ylo = sm.ylo[d], yhi = sm.yhi[d], xlo = sm.xlo[d], xhi = sm.xhi[d];
        VG_TR_ON(if (P.trace) tr_t = clock64();)
        // clear what this slice's Gaussian can read: the pairs overlapping the smoothed rows [gy0, gy1]
        // read buffer rows gy0 - 2 .. gy1 + 2, strips 1 .. ns + 2 (everything else is not looked at before
        // the normalised image overwrites it)
        {
            const int r_lo = max(max(ylo - 4, 0) + orow - 2, 0), r_hi = min(min(yhi + 2, Q - 1) + orow + 2, nr + 2);
            if (F_PITCH <= 32) {        // a lane per strip
                if (lane < ns + 3)
                    for (int r = r_lo + warp; r <= r_hi; r += NWF)
                        sm.B[r * F_PITCH + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
                for (int r = r_lo + warp; r <= r_hi; r += NWF)
                    for (int sb = lane; sb < ns + 3; sb += 32)
                        sm.B[r * F_PITCH + sb] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        __syncthreads();
        VG_TR_ON(if (P.trace) { const long long t = clock64(); tr_clear += t - tr_t; tr_t = t; })
        {
            // one item = one row of one point's 5x5 footprint (see project_image); all values are
            // positive floats: integer order == float order
            int *Bi = reinterpret_cast<int *>(sm.B) + orow * PW + ocol - 3 * PW - 3;
            const uint2 *cache = sm.cache + sm.base[d];
            if (xlo >= 3 && xhi <= Q - 2 && ylo >= 3 && yhi <= Q - 2) {      // no footprint leaves the image:
                if (STAMP == 1) {
                    // one item = one column of a footprint: the five lanes of a point hit five consecutive banks
                    for (int item = tid; item < 5 * cnt; item += NTF) {
                        const int p = (item * 13108) >> 16, dx = item - 5 * p;     // item / 5 for item < 10,240
                        const uint2 e = cache[p];
                        int *col = Bi + (int)(e.x >> 8) * PW + (int)(e.x & 255u) + dx;
#pragma unroll
                        for (int dy = 0; dy < 5; ++dy) atomicMax(col + dy * PW, (int)e.y);
                    }
                } else {
                    for (int p = tid; p < cnt; p += NTF) {                   // a thread stamps a whole footprint
                        const uint2 e = cache[p];
                        int *row = Bi + (int)(e.x >> 8) * PW + (int)(e.x & 255u);
#pragma unroll
                        for (int dy = 0; dy < 5; ++dy)
#pragma unroll
                            for (int dx = 0; dx < 5; ++dx) atomicMax(row + dy * PW + dx, (int)e.y);
                    }
                }
            } else {
                for (int item = tid; item < 5 * cnt; item += NTF) {
                    const int p = (item * 13108) >> 16, dy = item - 5 * p;
                    const uint2 e = cache[p];
                    const int X = (int)(e.x & 255u), Y = (int)(e.x >> 8);
                    const int q = Y - 3 + dy;
                    if (q < 0 || q > Q - 1) continue;
                    int *row = Bi + (Y + dy) * PW + X;
#pragma unroll
                    for (int dx = 0; dx < 5; ++dx) {
                        const int x = X - 3 + dx;
                        if (x >= 0 && x <= Q - 1) atomicMax(row + dx, (int)e.y);
                    }
                }
            }
        }
        __syncthreads();
        VG_TR_ON(if (P.trace) { const long long t = clock64(); tr_stamp += t - tr_t; tr_t = t; })
        {
            // smoothed rows of this slice: [gy0, gy1]; a pair takes part when one of its rows is inside
            const int gy0 = max(ylo - 4, 0), gy1 = min(yhi + 2, Q - 1);
            const unsigned lo = (unsigned)((gy0 + orow - 1) * PW), span = (unsigned)((gy1 - gy0 + 2) * PW);
            const float *Bf = reinterpret_cast<const float *>(sm.B);
#pragma unroll
            for (int k = 0; k < KQ; ++k) {
                if (k * NTF + (tid & ~31) >= npair) break;                 // warp-uniform
                const int o = boff[k] & 0x1fffffff;
                const bool act = tid + k * NTF < npair && (unsigned)(o - (int)lo) < span;
                if (!__any_sync(0xffffffffu, act)) continue;
                const float *cen = Bf + (act ? o : 4 * F_PITCH + 8);      // inactive lanes read a valid cell
                const bool ml = (boff[k] >> 29) & 1, mr = (boff[k] >> 30) & 1;
                float t[4][6];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const float *row = cen + (r - 1) * PW;
                    const float4 cc = *reinterpret_cast<const float4 *>(row);
                    float l = __shfl_up_sync(0xffffffffu, cc.w, 1);
                    float rr = __shfl_down_sync(0xffffffffu, cc.x, 1);
                    if (ml) l = row[-1];
                    if (mr) rr = row[4];
                    t[r][0] = l; t[r][1] = cc.x; t[r][2] = cc.y; t[r][3] = cc.z; t[r][4] = cc.w; t[r][5] = rr;
                }
                float o0[4], o1[4];
                gauss4(w, t[0], t[1], t[2], o0);
                gauss4(w, t[1], t[2], t[3], o1);
                if (act) {
                    img0[k] = max4(img0[k], make_float4(o0[0], o0[1], o0[2], o0[3]));
                    img1[k] = max4(img1[k], make_float4(o1[0], o1[1], o1[2], o1[3]));
                }
            }
        }
        VG_TR_ON(if (P.trace) tr_gauss += clock64() - tr_t;)
    }
    VG_TR(6);
    VG_TR_ON(if (P.trace && tid == 0 && b < kTraceImages) {
        P.trace[b * 12 + 9] = tr_clear; P.trace[b * 12 + 10] = tr_stamp; P.trace[b * 12 + 11] = tr_gauss;
    })

    // ---- phase 4: img / max(img), 1 - x; the normalised image goes back into B, margins = 1.0 ------
    {
        float m = 0.0f;
#pragma unroll
        for (int k = 0; k < KQ; ++k) {
            const int item = tid + k * NTF;
            if (item >= npair) {
                img0[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                img1[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                continue;
            }
            const int o = (boff[k] & 0x1fffffff) >> 2;
            const int rb = o / F_PITCH, sb = o - rb * F_PITCH;            // buffer row (first of the pair), strip
            if (rb + 1 > nr) img1[k] = make_float4(0.f, 0.f, 0.f, 0.f);   // second row of an odd last pair
            if (sb - 2 + us_lo == NS - 1) {                               // columns Q, Q + 1 are padding
                img0[k].z = 0.0f; img0[k].w = 0.0f;
                img1[k].z = 0.0f; img1[k].w = 0.0f;
            }
            m = fmaxf(fmaxf(m, fmaxf(img0[k].x, img0[k].y)), fmaxf(img0[k].z, img0[k].w));
            m = fmaxf(fmaxf(m, fmaxf(img1[k].x, img1[k].y)), fmaxf(img1[k].z, img1[k].w));
        }
        m = warp_max(m);
        if (lane == 0) sm.red[warp] = m;
        __syncthreads();                   // also: every Gaussian read of the last slice is done
        float mx = sm.red[0];
#pragma unroll
        for (int i = 1; i < NWF; ++i) mx = fmaxf(mx, sm.red[i]);
        const float rc_mx = __frcp_rn(mx);
        // 1 - t / mx with div_rn's three steps (q = t rc; r = t - mx q; q + r rc) on packed pairs: the same
        // IEEE operations per element as the scalar form
        const f32x2 rc2 = pack2(rc_mx, rc_mx), nmx2 = pack2(-mx, -mx), one2 = pack2(1.0f, 1.0f);
        auto norm2 = [&](float a, float b, float &ra, float &rb) {
            const f32x2 t = pack2(a, b);
            const f32x2 q = mul2(t, rc2);
            unpack2(sub2(one2, fma2(fma2(nmx2, q, t), rc2, q)), ra, rb);
        };
        auto norm4 = [&](float4 t) {
            norm2(t.x, t.y, t.x, t.y);
            norm2(t.z, t.w, t.z, t.w);
            return t;
        };
        const float4 one4 = make_float4(1.f, 1.f, 1.f, 1.f);
        if (TAB != 0) {
            // the halo rows 0, nr + 1, nr + 2 whole (a warp each); of the image rows only the four strips
            // beside the items that the emit can read -- 0, 1, ns + 2, ns + 3 (its source columns lie
            // within strips 0 .. ns + 3 for every (vlo, vhi): tests/test_host_side.py) -- item-wise
            static_assert(NWF >= 3, "one warp per halo row");
            if (warp < 3) {
                const int rb = warp == 0 ? 0 : nr + warp;
                for (int sb = lane; sb < F_PITCH; sb += 32) sm.B[rb * F_PITCH + sb] = one4;
            }
            for (int item = tid; item < 4 * nr; item += NTF) {
                const int rb = 1 + (item >> 2), e = item & 3;
                sm.B[rb * F_PITCH + (e < 2 ? e : ns + e)] = one4;
            }
        } else if (F_PITCH <= 32) {     // a lane per strip
            if (lane < F_PITCH) {
                const bool edge = lane < 2 || lane >= ns + 2;
                for (int rb = warp; rb < nr + 3; rb += NWF)
                    if (edge || rb == 0 || rb > nr) sm.B[rb * F_PITCH + lane] = one4;
            }
        } else {
            for (int sb = lane; sb < F_PITCH; sb += 32) {
                const bool edge = sb < 2 || sb >= ns + 2;
                for (int rb = warp; rb < nr + 3; rb += NWF)
                    if (edge || rb == 0 || rb > nr) sm.B[rb * F_PITCH + sb] = one4;
            }
        }
#pragma unroll
        for (int k = 0; k < KQ; ++k) {
            if (tid + k * NTF < npair) {
                const int o = (boff[k] & 0x1fffffff) >> 2;
                sm.B[o] = norm4(img0[k]);
                if (o / F_PITCH + 1 <= nr) sm.B[o + F_PITCH] = norm4(img1[k]);
            }
        }
        if (lane == 0) bulk_wait_read();   // the bulk stores have read their source long ago; no thread may
                                           // exit before that is certain
        __syncthreads();
        VG_TR(7);
        if (P.dbg_dens) {
            float *dd = P.dbg_dens + (size_t)b * Q * Q;
            const float *Bf = reinterpret_cast<const float *>(sm.B);
            for (int i = tid; i < Q * Q; i += NTF) {
                const int y = i / Q, x = i - y * Q;
                const int rb = y + orow, cb = x + ocol;
                const bool in = rb >= 0 && rb <= nr + 1 && cb >= 0 && cb < 4 * (ns + 4);
                dd[i] = in ? Bf[rb * PW + cb] : 1.0f;
            }
        }
    }

    // ---- phase 5: bilinear Q -> 224 (align_corners), floor(x*255), patch-major tiles ---------------
    // out = fma(HI[y0], lh0, HI[y0 + 1] * lh1), HI[y][ox] = fma(B[y][x0], lw0, B[y][x0 + 1] * lw1): the
    // contraction pattern of torch-CPU's separable interpolation.  (At the last source row / column the
    // reference clamps the second index; its weight is exactly 0 there and B holds a finite value one
    // step further, so the unclamped read gives the same +0 product.)
    if (oy_hi < oy_lo || g_hi < g_lo) return;
    {
        // a thread owns CW adjacent output columns and a run of SOURCE rows [ys, ye]: per source row one
        // horizontal interpolation (the row below is carried over in registers) and the two or three output
        // rows whose upper source row it is.  CW = 4: consecutive lanes read consecutive source pixels
        // (few bank conflicts: the shared-memory pipe is this kernel's busiest unit) and a row costs one
        // 8-byte store
        constexpr int CP = CW / 2;                                       // packed column pairs
        const int ncg = (g_hi - g_lo + 1) * (8 / CW);                    // column sets of the active groups
        const int nseg = small_div(NTF, ncg);
        const int seg = small_div(tid, ncg);
        if (seg >= nseg) return;
        const int ox0 = 8 * g_lo + CW * (tid - seg * ncg);
        const int ya = sm.i0[oy_lo], yb = sm.i0[oy_hi];
        const int len = small_div(yb - ya + nseg, nseg);
        const int ys = ya + seg * len, ye = min(ys + len - 1, yb);
        if (ys > ye) return;
        int oy = min((ys * (S - 1)) / (Q - 1), S - 1);                  // close to the first row with y0 == ys
        while (oy > oy_lo && (int)sm.i0[oy - 1] >= ys) --oy;
        while ((int)sm.i0[oy] < ys) ++oy;                                // i0 reaches yb >= ys at oy_hi
        oy = max(oy, oy_lo);
        // shared-window byte address of B[row 0][x0 of output column ox0 + j]: a source row is one add away
        const unsigned b0 = (unsigned)__cvta_generic_to_shared(sm.B) + 4u * (unsigned)(orow * PW + ocol);
        unsigned xa[CW];
        f32x2 lw0[CP], lw1[CP];
#pragma unroll
        for (int j = 0; j < CP; ++j) {
            xa[2 * j] = b0 + 4u * sm.i0[ox0 + 2 * j];
            xa[2 * j + 1] = b0 + 4u * sm.i0[ox0 + 2 * j + 1];
            if (TAB == 0) {
                const float2 ta = sm.lw[ox0 + 2 * j], tb = sm.lw[ox0 + 2 * j + 1];
                lw0[j] = pack2(ta.x, tb.x);
                lw1[j] = pack2(ta.y, tb.y);
            } else {        // the pairs as stored: one 8-byte load each, nothing to re-pack in the row loop
                lw0[j] = *reinterpret_cast<const f32x2 *>(&sm.cw0[(ox0 >> 1) + j]);
                lw1[j] = *reinterpret_cast<const f32x2 *>(&sm.cw1[(ox0 >> 1) + j]);
            }
        }
        f32x2 ha[CP], hb[CP];
        // TAB = 1, four columns per thread: their first source columns are s = i0[ox0] plus (0, d1, d2, d3) with
        // (d1, d2, d3) one of (0,0,1) (0,1,1) (1,1,1) (1,1,2) at R = 112 and (0,1,2) (1,2,3) at R = 224
        // (tests/test_host_side.py), so the thread loads the run s .. s + 3 (4) ONCE -- one address, half the
        // shared-memory wavefronts of eight scalar loads at stride 2 -- and picks its operands with
        // per-thread constant selectors.  The last float of the run may lie one past what the scalar form reads: still
        // inside the buffer row, never selected.
        // (byte-permute selectors rather than predicates: one instruction per pick, nothing to rematerialise)
        const unsigned s1 = sm.i0[ox0 + 1] != sm.i0[ox0] ? 0x7654u : 0x3210u,
                       s2 = sm.i0[ox0 + 2] != sm.i0[ox0] ? 0x7654u : 0x3210u,
                       s3 = (int)sm.i0[ox0 + 3] - (int)sm.i0[ox0] == 2 ? 0x7654u : 0x3210u;
        auto pick = [](unsigned sel, float a, float b) {       // sel = 0x3210: a, 0x7654: b
            return __uint_as_float(__byte_perm(__float_as_uint(a), __float_as_uint(b), sel));
        };
        auto hrow = [&](int y, f32x2 (&h)[CP]) {
            const unsigned ro = (unsigned)(y * (4 * PW));
            if (TAB != 0 && CW == 4 && R == 112) {
                float v0, v1, v2, v3;
                asm volatile("ld.shared.f32 %0, [%4];\n\tld.shared.f32 %1, [%4 + 4];\n\t"
                             "ld.shared.f32 %2, [%4 + 8];\n\tld.shared.f32 %3, [%4 + 12];"
                             : "=f"(v0), "=f"(v1), "=f"(v2), "=f"(v3) : "r"(xa[0] + ro));
                h[0] = fma2(pack2(v0, pick(s1, v0, v1)), lw0[0], mul2(pack2(v1, pick(s1, v1, v2)), lw1[0]));
                h[CP - 1] = fma2(pack2(pick(s2, v0, v1), pick(s3, v1, v2)), lw0[CP - 1],
                                 mul2(pack2(pick(s2, v1, v2), pick(s3, v2, v3)), lw1[CP - 1]));
                return;
            }
            if (TAB != 0 && CW == 4 && R == 224) {
                float v0, v1, v2, v3, v4;
                asm volatile("ld.shared.f32 %0, [%5];\n\tld.shared.f32 %1, [%5 + 4];\n\t"
                             "ld.shared.f32 %2, [%5 + 8];\n\tld.shared.f32 %3, [%5 + 12];\n\t"
                             "ld.shared.f32 %4, [%5 + 16];"
                             : "=f"(v0), "=f"(v1), "=f"(v2), "=f"(v3), "=f"(v4) : "r"(xa[0] + ro));
                h[0] = fma2(pack2(v0, pick(s1, v0, v1)), lw0[0], mul2(pack2(v1, pick(s1, v1, v2)), lw1[0]));
                h[CP - 1] = fma2(pack2(pick(s1, v1, v2), pick(s1, v2, v3)), lw0[CP - 1],
                                 mul2(pack2(pick(s1, v2, v3), pick(s1, v3, v4)), lw1[CP - 1]));
                return;
            }
#pragma unroll
            for (int j = 0; j < CP; ++j) {
                float a0, a1, c0, c1;
                asm volatile("ld.shared.f32 %0, [%2];\n\tld.shared.f32 %1, [%2 + 4];"
                             : "=f"(a0), "=f"(a1) : "r"(xa[2 * j] + ro));
                asm volatile("ld.shared.f32 %0, [%2];\n\tld.shared.f32 %1, [%2 + 4];"
                             : "=f"(c0), "=f"(c1) : "r"(xa[2 * j + 1] + ro));
                h[j] = fma2(pack2(a0, c0), lw0[j], mul2(pack2(a1, c1), lw1[j]));
            }
        };
        const f32x2 k255 = pack2(255.0f, 255.0f), kmagic = pack2(8388608.0f, 8388608.0f);
        op_t *tdst = tile ? tile + (ox0 >> 4) * 256 + (ox0 & 15) : nullptr;
        int ynext = sm.i0[oy];             // upper source row of output row oy, loaded one row ahead
        // the output rows whose upper source row is y, from top = HI[y] and bot = HI[y + 1]
        // `kTilesOnly` (a constant at both call sites once `walk` is inlined): the caller wants the tile and
        // no uint8 image -- the production case -- so the row loop carries neither pointer test (TAB = 1)
        auto emit_rows = [&](const bool kTilesOnly, int y, const f32x2 (&top)[CP], const f32x2 (&bot)[CP]) {
            while (oy <= oy_hi && ynext == y) {
                float2 lh;
                int toff = 0;               // byte offset of row oy inside the tile
                if (TAB == 0) {
                    lh = sm.lw[oy];
                    ynext = sm.i0[min(oy + 1, S - 1)];
                } else {
                    const float4 rec = sm.rowrec[oy];
                    lh = make_float2(rec.x, rec.y);
                    ynext = __float_as_int(rec.z);
                    toff = __float_as_int(rec.w);
                }
                const f32x2 h0 = pack2(lh.x, lh.x), h1 = pack2(lh.y, lh.y);
                unsigned fb[CW];
#pragma unroll
                for (int j = 0; j < CP; ++j) {
                    const f32x2 o = fma2(top[j], h0, mul2(bot[j], h1));
                    const f32x2 q = sub2(add2_rz(mul2(o, k255), kmagic), kmagic);
                    float q0, q1;
                    unpack2(q, q0, q1);
                    fb[2 * j] = __float_as_uint(q0);
                    fb[2 * j + 1] = __float_as_uint(q1);
                }
                if (kTilesOnly || tdst) {
                    unsigned pk[CP];
#pragma unroll
                    for (int j = 0; j < CP; ++j) {
#ifndef VG_OPERAND_BF16
                        pk[j] = pack_op(__uint_as_float(fb[2 * j]), __uint_as_float(fb[2 * j + 1]));
#else
                        pk[j] = __byte_perm(fb[2 * j], fb[2 * j + 1], 0x7632);   // integers 0..255: high halves
#endif
                    }
                    op_t *dst = TAB == 0 ? tdst + (oy >> 4) * (14 * 256) + (oy & 15) * 16
                                         : reinterpret_cast<op_t *>(reinterpret_cast<char *>(tdst) + toff);
                    if (CW == 8) *reinterpret_cast<uint4 *>(dst) = make_uint4(pk[0], pk[1 % CP], pk[2 % CP], pk[3 % CP]);
                    else if (CW == 4) *reinterpret_cast<uint2 *>(dst) = make_uint2(pk[0], pk[1 % CP]);
                    else *reinterpret_cast<unsigned *>(dst) = pk[0];
                }
                if (!kTilesOnly && u8) {
                    uint8_t *dst = u8 + oy * S + ox0;
                    unsigned w0 = 0, w1 = 0;
#pragma unroll
                    for (int j = 0; j < CW; ++j) {
                        const unsigned byte = (unsigned)__uint_as_float(fb[j]);
                        if (j < 4) w0 |= byte << (8 * j); else w1 |= byte << (8 * (j - 4));
                    }
                    if (CW == 8) *reinterpret_cast<uint2 *>(dst) = make_uint2(w0, w1);
                    else if (CW == 4) *reinterpret_cast<unsigned *>(dst) = w0;
                    else *reinterpret_cast<unsigned short *>(dst) = (unsigned short)w0;
                }
                ++oy;
            }
        };
        // two source rows per trip, the two row buffers swapping roles (no register copies)
        auto walk = [&](const bool tiles_only) {
            hrow(ys, ha);
            for (int y = ys;;) {
                hrow(y + 1, hb);
                emit_rows(tiles_only, y, ha, hb);
                if (++y > ye) break;
                hrow(y + 1, ha);
                emit_rows(tiles_only, y, hb, ha);
                if (++y > ye) break;
            }
        };
        if (TAB != 0 && tdst && !u8) walk(true);
        else walk(false);
        VG_TR(8);
    }
}

// Persistent CTAs (three per SM): the bilinear table and the constant background row are set up once,
// then images are drawn from a counter (they differ a lot in cost); the next index is fetched before the
// barrier that ends an image, so the loop adds one barrier per image and no exposed latency.
template <int R, int NTF, int MINB, int CW, int STAMP, int TAB>
__global__ void __launch_bounds__(NTF, MINB) projection_fast_kernel(const ProjParams P, const int images)
{
    static_assert(NTF >= S && NTF >= 64, "setup roles are mapped to thread ids");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FastSmem<R> &sm = *reinterpret_cast<FastSmem<R> *>(smem_raw);
    const int tid = threadIdx.x;
    if (tid < S) {
        const float l0 = __ldg(&P.tab->l0[tid]), l1 = __ldg(&P.tab->l1[tid]);
        sm.i0[tid] = (unsigned char)__ldg(&P.tab->i0[tid]);
        if (TAB == 0) {
            sm.lw[tid] = make_float2(l0, l1);
        } else {
            reinterpret_cast<float *>(sm.cw0)[tid] = l0;
            reinterpret_cast<float *>(sm.cw1)[tid] = l1;
            sm.rowrec[tid] = make_float4(l0, l1, __int_as_float(__ldg(&P.tab->i0[min(tid + 1, S - 1)])),
                                         __int_as_float(2 * ((tid >> 4) * (14 * 256) + (tid & 15) * 16)));
        }
    }
    if (P.bg_splat != 0u && P.tiles)
        for (int i = tid; i < 448; i += NTF) sm.bgsrc[i] = make_uint4(P.bg_splat, P.bg_splat, P.bg_splat, P.bg_splat);
    int i = (int)blockIdx.x;
    while (i < images) {
        fast_image<R, NTF, CW, STAMP, TAB>(P, P.block0 + i, sm);
        if (tid == 0) sm.next = (int)gridDim.x + atomicAdd(&P.defer[2], 1);
        __syncthreads();       // the image is finished by every warp; the next one re-initialises `sm`
        i = sm.next;
    }
}

template <int R, bool LIST>
int launch_projection_t(VgHandle *h, const ProjParams &P, long long blocks, cudaStream_t st)
{
    int rc = vg_set_smem_once(h, reinterpret_cast<const void *>(projection_kernel<R, LIST>), sizeof(Smem<R>));
    if (rc) return rc;
    projection_kernel<R, LIST><<<(unsigned)blocks, NT, sizeof(Smem<R>), st>>>(P);
    return VG_OK;
}

template <int R, int NTF, int MINB, int CW, int STAMP, int TAB>
int launch_fast_t(VgHandle *h, const ProjParams &P, long long blocks, cudaStream_t st)
{
    int rc = vg_set_smem_once(h, reinterpret_cast<const void *>(projection_fast_kernel<R, NTF, MINB, CW, STAMP, TAB>),
                              sizeof(FastSmem<R>));
    if (rc) return rc;
    const long long grid = std::min<long long>(blocks, (long long)MINB * h->num_sms);
    projection_fast_kernel<R, NTF, MINB, CW, STAMP, TAB><<<(unsigned)grid, NTF, sizeof(FastSmem<R>), st>>>(P, (int)blocks);
    return VG_OK;
}

// VG_PROJ_TRACE=1: mean cycles thread 0 of a CTA spends in each phase of the fast kernel (debug aid)
void print_trace(VgHandle *h, int images, cudaStream_t st)
{
    cudaStreamSynchronize(st);
    std::vector<long long> t((size_t)images * 12);
    cudaMemcpy(t.data(), h->proj_trace, t.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    const char *names[8] = {"phase 1 (rotate, min/max)", "quantise + atomics", "ranges, sort", "background",
                            "item setup", "slices", "normalise, write back", "emit (thread 0)"};
    double sum[8] = {0}, sl[3] = {0}, total = 0;
    int cnt = 0;
    for (int i = 0; i < images; ++i) {
        const long long *r = &t[(size_t)i * 12];
        if (r[0] == 0 || r[8] <= r[0]) continue;       // handed over / degenerate / idle thread 0
        for (int k = 0; k < 8; ++k) sum[k] += (double)(r[k + 1] - r[k]);
        for (int k = 0; k < 3; ++k) sl[k] += (double)r[9 + k];
        total += (double)(r[8] - r[0]);
        ++cnt;
    }
    if (!cnt) return;
    fprintf(stderr, "projection_fast_kernel trace: %d images, mean %.0f cycles per CTA\n", cnt, total / cnt);
    for (int k = 0; k < 8; ++k) fprintf(stderr, "  %-28s %8.0f  %5.1f %%\n", names[k], sum[k] / cnt, 100 * sum[k] / total);
    fprintf(stderr, "  slices: clear %.0f, stamp %.0f, gauss %.0f\n", sl[0] / cnt, sl[1] / cnt, sl[2] / cnt);
    cudaMemset(h->proj_trace, 0, t.size() * sizeof(long long));
}

constexpr long long kDeferCap = 1ll << 20;     // images per fast launch (capacity of the hand-over list)

}  // namespace

int projection_init(VgHandle *h)
{
    const int R = h->cfg.resolution, Q = R - 2;
    ProjTables *t = nullptr;
    VG_CUDA_CHECK(h, cudaMalloc(&t, sizeof(ProjTables)));
    h->proj_tables = t;
    projection_tables_kernel<<<64, 256>>>(t, Q);
    VG_CUDA_CHECK(h, cudaGetLastError());
    VG_CUDA_CHECK(h, cudaDeviceSynchronize());
    // point pools for clusters above CAP points: one per CTA that can be resident (~150 MB)
    int nsmid = 0;
    VG_CUDA_CHECK(h, cudaMemcpy(&nsmid, &t->nsmid, sizeof(int), cudaMemcpyDeviceToHost));
    if (nsmid < h->num_sms) nsmid = h->num_sms;
    h->proj_spill_sms = nsmid;
    const size_t slots = (size_t)nsmid * kSpillPerSm;
    VG_CUDA_CHECK(h, cudaMalloc(&h->proj_spill, slots * POOL * sizeof(uint2)));
    VG_CUDA_CHECK(h, cudaMalloc(&h->proj_spill_flags, slots * sizeof(int)));
    VG_CUDA_CHECK(h, cudaMemset(h->proj_spill_flags, 0, slots * sizeof(int)));
    if (R == 224)    // running depth-max image of the one CTA resident on each SM (~29 MB)
        VG_CUDA_CHECK(h, cudaMalloc(&h->proj_img_scratch, (size_t)nsmid * Q * R * sizeof(float)));
    // the fast kernel writes the background as a constant when the whole tile is one value (R = 112:
    // 255 everywhere), and derives the active output rows / columns from i0[o] = floor(o (Q-1) / (S-1))
    {
        std::vector<ProjTables> host(1);
        VG_CUDA_CHECK(h, cudaMemcpy(host.data(), t, sizeof(ProjTables), cudaMemcpyDeviceToHost));
        const uint16_t *bt = reinterpret_cast<const uint16_t *>(host[0].bg_tile);
        const uint8_t *bu = reinterpret_cast<const uint8_t *>(host[0].bg_u8);
        bool uniform = true;
        for (int i = 1; i < VG_TILE_ELEMS && uniform; ++i) uniform = bt[i] == bt[0];
        for (int i = 1; i < S * S && uniform; ++i) uniform = bu[i] == bu[0];
        h->proj_bg_splat = uniform && bt[0] != 0 ? ((uint32_t)bt[0] << 16) | bt[0] : 0u;
        h->proj_bg_u8_splat = 0x01010101u * bu[0];
        h->proj_table_exact = true;
        for (int o = 0; o < S; ++o)
            if (host[0].i0[o] != (o * (Q - 1)) / (S - 1)) h->proj_table_exact = false;
    }
#ifdef VG_PROJ_TRACE
    if (getenv("VG_PROJ_TRACE")) {
        VG_CUDA_CHECK(h, cudaMalloc(&h->proj_trace, (size_t)kTraceImages * 12 * sizeof(long long)));
        VG_CUDA_CHECK(h, cudaMemset(h->proj_trace, 0, (size_t)kTraceImages * 12 * sizeof(long long)));
    }
#endif
    // hand-over list of the fast kernel (count + image indices, 4 MB)
    VG_CUDA_CHECK(h, cudaMalloc(&h->proj_defer, (size_t)(kDeferCap + 3) * sizeof(int32_t)));
    // The fast kernel's shared-memory layout covers touched regions of F_MAXR rows x F_MAXS strips.
    // X, Y = clip(ceil(((u * obj_ratio + 1) / 2) * R), 1, R - 2) with |u| <= 1 up to a few ulps (1e-3 of
    // a cell covers them); the touched region adds 4 cells below and 2 above.  The kernel checks every
    // image against its layout anyway and hands over what does not fit.
    {
        const double rho = h->cfg.obj_ratio;
        const int lo = (int)std::max(1.0, std::ceil((1.0 - rho) * 0.5 * R - 1e-3));
        const int hi = (int)std::min((double)(R - 2), std::ceil((1.0 + rho) * 0.5 * R + 1e-3));
        const int rows = std::min(hi + 2, Q - 1) - std::max(lo - 4, 0) + 1;
        const int strips = (std::min(hi + 2, Q - 1) >> 2) - (std::max(lo - 4, 0) >> 2) + 1;
        const int max_rows = R == 112 ? FastGeo<112>::MAXR : FastGeo<224>::MAXR;
        const int max_strips = R == 112 ? FastGeo<112>::MAXS : FastGeo<224>::MAXS;
        h->proj_fast = rows <= max_rows && strips <= max_strips && h->proj_table_exact;
    }
    VG_CUDA_CHECK(h, cudaDeviceSynchronize());
    return VG_OK;
}

int launch_projection(VgHandle *h, const float *d_points, const int32_t *d_offsets, int32_t C,
                      op_t *d_tiles, uint8_t *d_u8, bool u8_first_only, int32_t *d_status,
                      const VgProjectDebug *dbg, cudaStream_t st)
{
    const VgConfig &cfg = h->cfg;
    if ((cfg.resolution != 112 && cfg.resolution != 224) || cfg.depth != D || cfg.image_size != S) {
        VG_SET_ERR(h, "projection kernel is specialised for R in {112, 224}, D=8, S=224 (got %d, %d, %d)",
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
    P.img_scratch = static_cast<float *>(h->proj_img_scratch);
    P.spill_sms = h->proj_spill_sms;
    P.C = C;
    P.V = cfg.num_views;
    memcpy(P.rot, cfg.rot, sizeof(P.rot));
    memcpy(P.gauss, cfg.gauss, sizeof(P.gauss));
    P.obj_ratio = (float)cfg.obj_ratio;
    P.depth_bias = (float)cfg.depth_bias;
    P.one_plus_bias = (float)(1.0 + cfg.depth_bias);
    P.rotate_mode = cfg.rotate_mode;
    P.div_mode = cfg.div_mode;
    P.tiles = d_tiles;
    P.u8 = d_u8;
    P.u8_first_only = u8_first_only ? 1 : 0;
    P.status = d_status;
    P.dbg_grid = dbg ? dbg->d_grid : nullptr;
    P.dbg_dens = dbg ? dbg->d_densified : nullptr;
    P.defer = static_cast<int32_t *>(h->proj_defer);
    P.block0 = 0;
    P.trace = h->proj_trace;
    P.bg_splat = h->proj_bg_splat;
    P.bg_u8_splat = h->proj_bg_u8_splat;
    const long long blocks = (long long)C * cfg.num_views;
    // algorithmic bytes recorded here: the emitted tiles; the caller adds 12 * sum(N) for the points
    VgProfScope prof(h, VG_K_PROJECTION, (double)blocks * VG_TILE_ELEMS * 2.0, st);
    int rc;
    if (h->proj_fast && h->sw.proj_variant != 0 && !P.dbg_grid) {
        // fast kernel over every image; what it hands over (clusters above FAST_N points) runs in the
        // general kernel in list mode, two CTAs per SM striding over the list
        for (long long b0 = 0; b0 < blocks; b0 += kDeferCap) {
            const long long nb = std::min(kDeferCap, blocks - b0);
            VG_CUDA_CHECK(h, cudaMemsetAsync(P.defer, 0, 3 * sizeof(int32_t), st));
            P.block0 = (int32_t)b0;
            if (cfg.resolution == 224) {
                switch (h->sw.proj_variant) {
                case 2: rc = launch_fast_t<224, 512, 1, 4, 1, 0>(h, P, nb, st); break;
                default: rc = launch_fast_t<224, 512, 1, 4, 1, 1>(h, P, nb, st); break;
                }
            } else {
                switch (h->sw.proj_variant) {
                case 2: rc = launch_fast_t<112, 256, 3, 4, 1, 0>(h, P, nb, st); break;
                default: rc = launch_fast_t<112, 256, 3, 4, 1, 1>(h, P, nb, st); break;
                }
            }
            if (rc) return rc;
            VG_LAUNCH_CHECK(h);
            rc = cfg.resolution == 224
                     ? launch_projection_t<224, true>(h, P, std::min<long long>(nb, (long long)h->num_sms), st)
                     : launch_projection_t<112, true>(h, P, std::min<long long>(nb, 2ll * h->num_sms), st);
            if (rc) return rc;
            VG_LAUNCH_CHECK(h);
        }
        if (P.trace) print_trace(h, (int)std::min<long long>(blocks, kTraceImages), st);
        return VG_OK;
    }
    rc = cfg.resolution == 112 ? launch_projection_t<112, false>(h, P, blocks, st)
                               : launch_projection_t<224, false>(h, P, blocks, st);
    if (rc) return rc;
    VG_LAUNCH_CHECK(h);
    return VG_OK;
}

}  // namespace vg
