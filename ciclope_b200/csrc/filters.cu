// Gaussian smoothing ahead of the segmentation (pipeline.py:146-148: skimage.filters.gaussian(I, sigma,
// preserve_range=True) = scipy.ndimage.gaussian_filter(I.astype(float64), sigma, mode='nearest',
// truncate=4.0)) and the threshold of the resulting float64 image (preprocess.py:46).
//
// scipy filters one axis after the other (0, 1, 2), each pass in float64 with the symmetric-kernel loop of
// NI_Correlate1D:   o = x[0] * w[0];  for j = -r .. -1:  o += (x[j] + x[-j]) * w[j]
// (farthest pair first, no fused multiply-add, 'nearest' = clamped indices).  The kernels below do exactly
// that arithmetic with __dmul_rn / __dadd_rn, so the result has scipy's bits.
#include <cstdint>

#include "common.cuh"

#define GAUSS_MAX_RADIUS 255

template <typename TIN>
__global__ void __launch_bounds__(256)
k_gauss1d(const TIN* __restrict__ in, double* __restrict__ out, i64 n_total, i64 n_axis, i64 stride,
          const double* __restrict__ weights, int r)
{
    __shared__ double w[2 * GAUSS_MAX_RADIUS + 1];
    for (int i = threadIdx.x; i < 2 * r + 1; i += blockDim.x) w[i] = weights[i];
    __syncthreads();
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n_total; i += (i64)gridDim.x * blockDim.x) {
        const i64 pos = (stride == 1 && n_total < (1LL << 32)) ? (i64)((u32)i % (u32)n_axis)
                                                               : (i64)(((u64)i / (u64)stride) % (u64)n_axis);
        const TIN* line = in + (i - pos * stride);
        double acc = __dmul_rn((double)line[pos * stride], w[r]);
        for (int j = -r; j < 0; ++j) {
            i64 a = pos + j, b = pos - j;
            if (a < 0) a = 0;
            if (b > n_axis - 1) b = n_axis - 1;
            const double s = __dadd_rn((double)line[a * stride], (double)line[b * stride]);
            acc = __dadd_rn(acc, __dmul_rn(s, w[r + j]));
        }
        out[i] = acc;
    }
}

// Strided axes (inner > 1): the array is [outer][n_axis][inner] with `inner` contiguous elements.  A CTA
// owns a compact tile of 32 inner x 64 axis positions (thread = 8 consecutive axis positions of one
// column), so the 2 r + 1 taps of neighbouring outputs are re-read from L1 instead of L2.
#define GAUSS_TA 64
template <typename TIN>
__global__ void __launch_bounds__(256)
k_gauss_strided(const TIN* __restrict__ in, double* __restrict__ out, i64 n_axis, i64 inner, u32 tiles_inner,
                const double* __restrict__ weights, int r)
{
    __shared__ double w[2 * GAUSS_MAX_RADIUS + 1];
    const int tid = threadIdx.y * 32 + threadIdx.x;
    for (int i = tid; i < 2 * r + 1; i += 256) w[i] = weights[i];
    __syncthreads();
    const u32 ba = blockIdx.x / tiles_inner;
    const u32 bx = blockIdx.x - ba * tiles_inner;
    const i64 col = (i64)bx * 32 + threadIdx.x;
    if (col >= inner) return;
    const i64 a0 = (i64)ba * GAUSS_TA + threadIdx.y * (GAUSS_TA / 8);
    const i64 base = (i64)blockIdx.y * n_axis * inner + col;
    const TIN* line = in + base;
    if (a0 - r >= 0 && a0 + GAUSS_TA / 8 - 1 + r <= n_axis - 1) {
        // interior: no clamping, running offsets instead of an index multiply per tap
        const i64 span = (i64)r * inner;
#pragma unroll 2
        for (int t = 0; t < GAUSS_TA / 8; ++t) {
            const TIN* c = line + (a0 + t) * inner;
            double acc = __dmul_rn((double)c[0], w[r]);
            i64 off = span;
            for (int j = 0; j < r; ++j, off -= inner)
                acc = __dadd_rn(acc, __dmul_rn(__dadd_rn((double)c[-off], (double)c[off]), w[j]));
            out[base + (a0 + t) * inner] = acc;
        }
        return;
    }
#pragma unroll 1
    for (int t = 0; t < GAUSS_TA / 8; ++t) {
        const i64 pos = a0 + t;
        if (pos >= n_axis) break;
        double acc = __dmul_rn((double)line[pos * inner], w[r]);
        for (int j = -r; j < 0; ++j) {
            i64 a = pos + j, b = pos - j;
            if (a < 0) a = 0;
            if (b > n_axis - 1) b = n_axis - 1;
            const double s = __dadd_rn((double)line[a * inner], (double)line[b * inner]);
            acc = __dadd_rn(acc, __dmul_rn(s, w[r + j]));
        }
        out[base + pos * inner] = acc;
    }
}

// Contiguous axis (stride 1): a CTA stages GAUSS_TX outputs' worth of one line (+ the clamped halo) in shared
// memory as float64, then every thread reads its 2 r + 1 taps from there.
#define GAUSS_TX 1024
template <typename TIN>
__global__ void __launch_bounds__(256)
k_gauss_rows(const TIN* __restrict__ in, double* __restrict__ out, i64 n_axis, u32 chunks,
             const double* __restrict__ weights, int r)
{
    __shared__ double w[2 * GAUSS_MAX_RADIUS + 1];
    __shared__ double sx[GAUSS_TX + 2 * GAUSS_MAX_RADIUS];
    for (int i = threadIdx.x; i < 2 * r + 1; i += 256) w[i] = weights[i];
    const i64 line = (i64)(blockIdx.x / chunks);
    const i64 x0 = (i64)(blockIdx.x - (u32)line * chunks) * GAUSS_TX;
    const TIN* src = in + line * n_axis;
    const int span = (int)((n_axis - x0 < GAUSS_TX ? n_axis - x0 : GAUSS_TX)) + 2 * r;
    for (int i = threadIdx.x; i < span; i += 256) {
        i64 x = x0 - r + i;
        if (x < 0) x = 0;
        if (x > n_axis - 1) x = n_axis - 1;
        sx[i] = (double)src[x];
    }
    __syncthreads();
    double* dst = out + line * n_axis + x0;
    for (int o = threadIdx.x; o < span - 2 * r; o += 256) {
        const double* c = sx + o + r;
        double acc = __dmul_rn(c[0], w[r]);
        for (int j = -r; j < 0; ++j) acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(c[j], c[-j]), w[r + j]));
        dst[o] = acc;
    }
}

static int gauss_grid(i64 n_total)
{
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    i64 blocks = (n_total + 255) / 256;
    if (blocks > (i64)sms * 16) blocks = (i64)sms * 16;
    return (int)blocks;
}

// One pass of scipy.ndimage.correlate1d(input, weights, axis, mode='nearest') for a symmetric kernel of
// 2 r + 1 weights over an array of n_total elements whose filtered axis has n_axis entries `stride` elements
// apart (C order: stride = product of the later dimensions).  in_is_u8: the input is uint8 (first pass),
// else float64.  in and out must not alias.
extern "C" int cfe_gauss1d(const void* in, int in_is_u8, double* out, int64_t n_total, int64_t n_axis,
                           int64_t stride, const double* weights, int r, cudaStream_t stream)
{
    if (n_total <= 0) return 0;
    if (r < 0 || r > GAUSS_MAX_RADIUS) { cfe_set_error("cfe_gauss1d: radius must be in 0..%d", GAUSS_MAX_RADIUS); return -1; }
    if (n_axis <= 0 || stride <= 0 || in == (const void*)out || n_total % (n_axis * stride)) {
        cfe_set_error("cfe_gauss1d: bad arguments");
        return -1;
    }
    const i64 outer = n_total / (n_axis * stride);
    const i64 tiles_inner = (stride + 31) / 32, tiles_axis = (n_axis + GAUSS_TA - 1) / GAUSS_TA;
    if (stride >= 32 && outer <= 65535 && tiles_inner * tiles_axis < (1LL << 31)) {
        const dim3 grid((unsigned)(tiles_inner * tiles_axis), (unsigned)outer), block(32, 8);
        if (in_is_u8)
            k_gauss_strided<uint8_t><<<grid, block, 0, stream>>>(reinterpret_cast<const uint8_t*>(in), out, n_axis,
                                                                 stride, (u32)tiles_inner, weights, r);
        else
            k_gauss_strided<double><<<grid, block, 0, stream>>>(reinterpret_cast<const double*>(in), out, n_axis,
                                                                stride, (u32)tiles_inner, weights, r);
        CFE_CHECK_LAUNCH("k_gauss_strided");
        return 0;
    }
    const i64 chunks = (n_axis + GAUSS_TX - 1) / GAUSS_TX;
    if (stride == 1 && n_axis >= 64 && (n_total / n_axis) * chunks < (1LL << 31)) {
        const unsigned grid = (unsigned)((n_total / n_axis) * chunks);
        if (in_is_u8)
            k_gauss_rows<uint8_t><<<grid, 256, 0, stream>>>(reinterpret_cast<const uint8_t*>(in), out, n_axis,
                                                           (u32)chunks, weights, r);
        else
            k_gauss_rows<double><<<grid, 256, 0, stream>>>(reinterpret_cast<const double*>(in), out, n_axis,
                                                          (u32)chunks, weights, r);
        CFE_CHECK_LAUNCH("k_gauss_rows");
        return 0;
    }
    const int blocks = gauss_grid(n_total);
    if (in_is_u8)
        k_gauss1d<uint8_t><<<blocks, 256, 0, stream>>>(reinterpret_cast<const uint8_t*>(in), out, n_total, n_axis,
                                                      stride, weights, r);
    else
        k_gauss1d<double><<<blocks, 256, 0, stream>>>(reinterpret_cast<const double*>(in), out, n_total, n_axis,
                                                     stride, weights, r);
    CFE_CHECK_LAUNCH("k_gauss1d");
    return 0;
}

__global__ void __launch_bounds__(256)
k_threshold_f64(const double* __restrict__ img, i64 n, double t, uint8_t* __restrict__ out)
{
    // 4 values per thread -> one 32-bit store
    const i64 n4 = n >> 2;
    for (i64 g = (i64)blockIdx.x * blockDim.x + threadIdx.x; g < n4; g += (i64)gridDim.x * blockDim.x) {
        const double2 a = __ldg(reinterpret_cast<const double2*>(img) + 2 * g);
        const double2 b = __ldg(reinterpret_cast<const double2*>(img) + 2 * g + 1);
        const u32 r = (a.x > t ? 1u : 0u) | (a.y > t ? 0x100u : 0u) | (b.x > t ? 0x10000u : 0u) | (b.y > t ? 0x1000000u : 0u);
        reinterpret_cast<u32*>(out)[g] = r;
    }
    for (i64 i = (n4 << 2) + (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x)
        out[i] = img[i] > t ? 1 : 0;
}

// out[i] = img[i] > t for a float64 image (NaN compares false, as in NumPy); img 16-byte, out 4-byte aligned
extern "C" int cfe_threshold_f64(const double* img, int64_t n, double t, uint8_t* out, cudaStream_t stream)
{
    if (n <= 0) return 0;
    if ((reinterpret_cast<uintptr_t>(img) & 15u) || (reinterpret_cast<uintptr_t>(out) & 3u)) {
        cfe_set_error("cfe_threshold_f64: misaligned buffers");
        return -1;
    }
    k_threshold_f64<<<gauss_grid(n / 4 + 1), 256, 0, stream>>>(img, n, t, out);
    CFE_CHECK_LAUNCH("k_threshold_f64");
    return 0;
}

// ------------------------------------------------------------------------------------------
// Spline resampling (preprocess.py:49-75: scipy.ndimage.zoom(image, 1 / factor, order=2)).
//
// scipy: (1) quadratic B-spline pre-filter along every axis in float64 -- gain (1 - z)(1 - 1/z), causal
// recursion c[i] += z c[i-1] started from the mirror-boundary sum, anti-causal recursion
// c[i] = z (c[i+1] - c[i]), pole z = sqrt(8) - 3 (ni_splines.c); (2) every output sample is the sum of the
// 3 x 3 x 3 coefficients around its input coordinate o * (n_in - 1) / (n_out - 1) weighted by the B-spline
// (taps beyond the edge mirror; a coordinate outside [0, n_in - 1] gives the constant 0).
// Parity: NOT bit-exact (scipy's compiled recursion is not reproduced to the last bit by any evaluation
// order tried); the stated tolerance is 1e-12 absolute on float64 images of 8-bit range, and identical
// uint8 / thresholded results on the test volumes (tests/golden/resample.npz).
//
// k_spline_prefilter: one thread per line, consecutive threads = consecutive positions of the fastest axis
// that is not filtered (coalesced for the z and y passes).  The mirror sum is cut after 96 terms
// (|z|^96 < 1e-73).
#define SPLINE_HORIZON 96

__global__ void __launch_bounds__(128)
k_spline_prefilter(double* __restrict__ data, i64 n_lines, i64 n, i64 stride, i64 inner, double z)
{
    const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_lines || n < 2) return;
    // line t: outer index t / inner, inner index t % inner  ->  base = outer * n * inner + in
    const i64 outer = t / inner, in = t - outer * inner;
    double* c = data + outer * n * inner + in;
    const double gain = (1.0 - z) * (1.0 - 1.0 / z);
    const double z_n_1 = pow(z, (double)(n - 1));
    const i64 hor = n - 1 < SPLINE_HORIZON ? n - 1 : SPLINE_HORIZON;
    double c0 = c[0] * gain + z_n_1 * (c[(n - 1) * stride] * gain);
    double z_i = z;
    for (i64 i = 1; i < hor; ++i) {
        c0 += z_i * (c[i * stride] * gain + z_n_1 * (c[(n - 1 - i) * stride] * gain));
        z_i *= z;
    }
    double prev = c0 / (1.0 - z_n_1 * z_n_1);
    c[0] = prev;
    for (i64 i = 1; i < n; ++i) {
        prev = c[i * stride] * gain + z * prev;
        c[i * stride] = prev;
    }
    double nxt = (z * c[(n - 2) * stride] + c[(n - 1) * stride]) * z / (z * z - 1.0);
    c[(n - 1) * stride] = nxt;
    for (i64 i = n - 2; i >= 0; --i) {
        nxt = z * (nxt - c[i * stride]);
        c[i * stride] = nxt;
    }
}

extern "C" int cfe_spline_prefilter(double* data, int64_t outer, int64_t n, int64_t inner, cudaStream_t stream)
{
    if (outer <= 0 || n <= 0 || inner <= 0) return 0;
    const i64 n_lines = outer * inner;
    k_spline_prefilter<<<(unsigned)((n_lines + 127) / 128), 128, 0, stream>>>(data, n_lines, n, inner, inner,
                                                                             sqrt(8.0) - 3.0);
    CFE_CHECK_LAUNCH("k_spline_prefilter");
    return 0;
}

// Interpolation: tap tables per axis from the host (idx [n_out][3] input indices, w [n_out][3] weights,
// zero [n_out] flags), one thread per output sample, 27 taps.  OUT_U8: scipy's conversion of the float64 sum
// to uint8 (t > 0 ? t + 0.5 : 0, clamped to 255, truncated).
template <bool OUT_U8>
__global__ void __launch_bounds__(256)
k_spline_zoom(const double* __restrict__ coef, i64 Yi, i64 Xi, i64 Zo, i64 Yo, i64 Xo,
              const int* __restrict__ iz, const int* __restrict__ iy, const int* __restrict__ ix,
              const double* __restrict__ wz, const double* __restrict__ wy, const double* __restrict__ wx,
              const uint8_t* __restrict__ zz, const uint8_t* __restrict__ zy, const uint8_t* __restrict__ zx,
              void* __restrict__ out)
{
    const i64 o = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= Zo * Yo * Xo) return;
    const i64 s = o / (Yo * Xo), rem = o - s * Yo * Xo, r = rem / Xo, c = rem - r * Xo;
    double t = 0.0;
    if (!(zz[s] | zy[r] | zx[c])) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double* pz = coef + (i64)iz[3 * s + a] * Yi * Xi;
            const double wa = wz[3 * s + a];
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                const double* py = pz + (i64)iy[3 * r + b] * Xi;
                const double wab = wa * wy[3 * r + b];
#pragma unroll
                for (int d = 0; d < 3; ++d) t += py[ix[3 * c + d]] * (wab * wx[3 * c + d]);
            }
        }
    }
    if (OUT_U8) {
        double v = t > 0.0 ? t + 0.5 : 0.0;
        if (v > 255.0) v = 255.0;
        reinterpret_cast<uint8_t*>(out)[o] = (uint8_t)v;
    } else {
        reinterpret_cast<double*>(out)[o] = t;
    }
}

extern "C" int cfe_spline_zoom(const double* coef, int64_t Yi, int64_t Xi, int64_t Zo, int64_t Yo, int64_t Xo,
                               const int* iz, const int* iy, const int* ix, const double* wz, const double* wy,
                               const double* wx, const uint8_t* zz, const uint8_t* zy, const uint8_t* zx,
                               int out_u8, void* out, cudaStream_t stream)
{
    const i64 n = Zo * Yo * Xo;
    if (n <= 0) return 0;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    if (out_u8)
        k_spline_zoom<true><<<blocks, 256, 0, stream>>>(coef, Yi, Xi, Zo, Yo, Xo, iz, iy, ix, wz, wy, wx, zz, zy, zx, out);
    else
        k_spline_zoom<false><<<blocks, 256, 0, stream>>>(coef, Yi, Xi, Zo, Yo, Xo, iz, iy, ix, wz, wy, wx, zz, zy, zx, out);
    CFE_CHECK_LAUNCH("k_spline_zoom");
    return 0;
}
