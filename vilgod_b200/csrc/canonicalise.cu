// Cluster canonicalisation on the GPU (SURVEY.md section 8 row a0 / "next" row f1): the per-cluster
// host loop of ZeroShotDetector.classification before the projection,
//   apply_transform(pts, transform_to_ego)            src/utils/pointcloud_utils.py:21-46
//   transform_cluster_points_to_origin(pts)           src/utils/pointcloud_utils.py:390-412
// call site src/vilgod/zero_shot_detector.py:391-394, for a whole packed frame in one launch.
//
// One CTA per cluster.  Same dtype at every step as the reference: ego points are rounded to
// fp32 (the reference writes them back into an fp32 array), the xy medians are numpy's fp32
// medians (mean of the two middle order statistics), the yaw angle is an fp32 atan2, everything
// after is float64, and the result is cast to fp32 once.  The medians come from an exact radix
// select on order-preserving integer keys, so they are bit-identical to numpy's; the float64 chain
// differs from scipy's quaternion round trip by ~1e-16 relative and the fp32 atan2 of the yaw by up
// to ~2 ulp from glibc's, i.e. coordinates match the host path to <= 2e-6 m (tests), most bit-equal.
#include "common.cuh"

namespace vg {
namespace {

constexpr int CT = 256;

__device__ __forceinline__ unsigned f2key(float f)
{
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(unsigned k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// ego-frame fp32 coordinate `axis` of point i: fp32( T[axis,:3] . p + T[axis,3] ) in float64
__device__ __forceinline__ float ego_coord(const float *__restrict__ p, const double *T, int axis)
{
    if (!T) return p[axis];
    const double v = T[4 * axis + 0] * (double)p[0] + T[4 * axis + 1] * (double)p[1] +
                     T[4 * axis + 2] * (double)p[2] + T[4 * axis + 3];
    return (float)v;
}

// k-th smallest (0-based) of the ego coordinate `axis` over the cluster: 4-pass 8-bit radix select
__device__ float radix_select(const float *__restrict__ pts, int n, const double *T, int axis, int k,
                              unsigned *hist /*[256] smem*/, unsigned *sh /*[2] smem*/)
{
    unsigned prefix = 0, mask = 0;
    int kk = k;
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        for (int i = threadIdx.x; i < 256; i += CT) hist[i] = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += CT) {
            const unsigned key = f2key(ego_coord(pts + 3 * (size_t)i, T, axis));
            if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 0xffu], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned acc = 0;
            int b = 0;
            for (; b < 256; ++b) {
                if (acc + hist[b] > (unsigned)kk) break;
                acc += hist[b];
            }
            sh[0] = (unsigned)b;
            sh[1] = acc;
        }
        __syncthreads();
        prefix |= sh[0] << shift;
        mask |= 0xffu << shift;
        kk -= (int)sh[1];
        __syncthreads();
    }
    return key2f(prefix);
}

__global__ void __launch_bounds__(CT) canonicalise_kernel(const float *__restrict__ pts_in,
                                                          const int32_t *__restrict__ offsets,
                                                          const double *__restrict__ T_ego,
                                                          float *__restrict__ pts_out,
                                                          int32_t *__restrict__ status)
{
    __shared__ unsigned hist[256];
    __shared__ unsigned sh[2];
    __shared__ double tt[16];
    const int c = blockIdx.x;
    const int beg = offsets[c], n = offsets[c + 1] - beg;
    if (n <= 0) {
        if (status && threadIdx.x == 0) status[c] = VG_EDEGENERATE;
        return;
    }
    if (status && threadIdx.x == 0) status[c] = VG_OK;
    const double *T = nullptr;
    if (T_ego) {
        if (threadIdx.x < 16) tt[threadIdx.x] = T_ego[threadIdx.x];
        __syncthreads();
        T = tt;
    }
    const float *p = pts_in + 3 * (size_t)beg;
    // numpy.median of an fp32 column: mean of the two middle order statistics, in fp32
    const int k0 = (n - 1) >> 1, k1 = n >> 1;
    float cx = radix_select(p, n, T, 0, k0, hist, sh);
    float cy = radix_select(p, n, T, 1, k0, hist, sh);
    if (k1 != k0) {
        const float cx1 = radix_select(p, n, T, 0, k1, hist, sh);
        const float cy1 = radix_select(p, n, T, 1, k1, hist, sh);
        cx = __fmul_rn(__fadd_rn(cx, cx1), 0.5f);
        cy = __fmul_rn(__fadd_rn(cy, cy1), 0.5f);
    }
    // np.arctan2 on fp32 scalars.  CUDA's atan2f and glibc's are both accurate to ~1-2 ulp but not
    // identical: the yaw may differ in its last bits, which moves coordinates by <= ~1e-6 m (the
    // reason this stage's parity is tolerance based)
    const float angle = atan2f(cy, cx);
    const double a = -(double)angle;
    const double ca = cos(a), sa = sin(a);
    // rot = Rx(pi) @ Rz(pi/2) exactly as float64 evaluates it (cos(pi/2) = 6.1e-17, not 0)
    const double c2 = cos(1.5707963267948966), s2 = sin(1.5707963267948966);
    const double cp = cos(3.141592653589793), sp = sin(3.141592653589793);
    // Rz(pi/2) = [[c2,-s2,0],[s2,c2,0],[0,0,1]],  Rx(pi) = [[1,0,0],[0,cp,-sp],[0,sp,cp]]
    const double m00 = c2, m01 = -s2, m02 = 0.0;
    const double m10 = cp * s2, m11 = cp * c2, m12 = -sp;
    const double m20 = sp * s2, m21 = sp * c2, m22 = cp;
    for (int i = threadIdx.x; i < n; i += CT) {
        const float *q = p + 3 * (size_t)i;
        const float ex = ego_coord(q, T, 0), ey = ego_coord(q, T, 1), ez = ego_coord(q, T, 2);
        const double x = (double)__fsub_rn(ex, cx), y = (double)__fsub_rn(ey, cy), z = (double)ez;
        double rx = ca * x - sa * y;            // rotate by -angle about z
        const double ry = sa * x + ca * y;
        rx -= 1.0;                              // shift one metre into x
        const double vx = z, vy = ry, vz = rx;  // reorder to (z, y, x)
        float *o = pts_out + 3 * ((size_t)beg + i);
        o[0] = (float)(m00 * vx + m01 * vy + m02 * vz);
        o[1] = (float)(m10 * vx + m11 * vy + m12 * vz);
        o[2] = (float)(m20 * vx + m21 * vy + m22 * vz);
    }
}

}  // namespace

int launch_canonicalise(VgHandle *h, const float *d_in, const int32_t *d_offsets, int32_t C,
                        const double *d_transform, float *d_out, int32_t *d_status, cudaStream_t st)
{
    if (C <= 0) return VG_OK;
    canonicalise_kernel<<<(unsigned)C, CT, 0, st>>>(d_in, d_offsets, d_transform, d_out, d_status);
    VG_LAUNCH_CHECK(h);
    return VG_OK;
}

}  // namespace vg
