// Slice-wise binary morphology on packed bits for periosteummask (preprocess.py:357-420):
//   skimage.morphology.binary_closing(slice, disk(r))  =  erosion(dilation(slice))  with
//   scipy.ndimage.binary_dilation(border_value = 0) and binary_erosion(border_value = 1), disk(r) = {x^2 + y^2 <= r^2};
//   scipy.ndimage.binary_fill_holes(slice)  =  the background that is not 4-connected to the slice border.
//
// Dilation by a disk, word-parallel:  out = OR_{dy} OR_{|dx| <= w(dy)} row[y+dy] << dx  is regrouped by dx:
// out = OR_k ( V_k << k  |  V_k >> k ),  V_k = OR of the rows |dy| <= h(k),  h(k) = floor(sqrt(r^2 - k^2)).  V_k grows
// as k falls, so every one of the 2r + 1 rows is read once and every k costs two funnel shifts: ~130
// instructions per 32 output pixels at r = 10 instead of ~840 for the row-by-row form.
// Erosion with border value 1 is the complement of the dilation of the complement (inside the image), so the
// same kernel does both: `invert_in` complements the valid input bits.  The background of the closed slice --
// what fill_holes floods -- is dilation(~dilation(x)) directly, no final complement needed.
#include <cstdint>

#include "common.cuh"

struct DiskParams {
    int r;
    int h[32];          // h[k] = largest |dy| with k^2 + dy^2 <= r^2
};

__global__ void __launch_bounds__(256)
k_disk_dilate_bits(const u32* __restrict__ in, i64 n_words, int W, FastDiv dW, FastDiv dY, u32 Y, int X,
                   int invert_in, DiskParams P, u32* __restrict__ out)
{
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_words) return;
    const u32 row = fd_div((u32)i, dW);
    const int w = (int)((u32)i - row * (u32)W);
    const u32 z = fd_div(row, dY);
    const int y = (int)(row - z * Y);
    const u32 tail = (X & 31) ? ((1u << (X & 31)) - 1u) : 0xffffffffu;      // valid bits of the last word
    auto load = [&](int yy, int ww) -> u32 {
        if (yy < 0 || yy >= (int)Y || ww < 0 || ww >= W) return 0u;
        u32 v = __ldg(in + ((i64)z * Y + yy) * W + ww);
        if (invert_in) v = ~v;
        if (ww == W - 1) v &= tail;
        return v;
    };
    u32 L = 0, M = 0, R = 0, acc = 0;
    int have = -1;                                 // rows |dy| <= have are in (L, M, R)
    for (int k = P.r; k >= 0; --k) {
        const int hk = P.h[k];
        for (int dy = have + 1; dy <= hk; ++dy) {
            L |= load(y - dy, w - 1); M |= load(y - dy, w); R |= load(y - dy, w + 1);
            if (dy) { L |= load(y + dy, w - 1); M |= load(y + dy, w); R |= load(y + dy, w + 1); }
        }
        if (hk > have) have = hk;
        if (k == 0) acc |= M;
        else acc |= __funnelshift_l(L, M, k) | __funnelshift_r(M, R, k);
    }
    if (w == W - 1) acc &= tail;
    out[i] = acc;
}

// out = dilation of every (Y, X) slice of `in` by disk(r), 0 outside the slice; invert_in: dilate the complement
// (erosion with border value 1 = ~ of that).  bits layout of cfe_pack_bits: [Z*Y rows][ceil(X/32) words].
extern "C" int cfe_disk_dilate_bits(const uint32_t* in, int64_t Z, int64_t Y, int64_t X, int r, int invert_in,
                                    uint32_t* out, cudaStream_t stream)
{
    if (Z <= 0 || Y <= 0 || X <= 0) return 0;
    if (r < 0 || r > 31) { cfe_set_error("cfe_disk_dilate_bits: radius must be in 0..31"); return -1; }
    const int W = (int)((X + 31) / 32);
    const i64 n_words = Z * Y * W;
    if (n_words >= (1LL << 32) || in == out) { cfe_set_error("cfe_disk_dilate_bits: bad arguments"); return -1; }
    DiskParams P;
    P.r = r;
    for (int k = 0; k < 32; ++k) {
        int h = -1;
        if (k <= r) { h = 0; while ((h + 1) * (h + 1) + k * k <= r * r) ++h; }
        P.h[k] = h;
    }
    k_disk_dilate_bits<<<(unsigned)((n_words + 255) / 256), 256, 0, stream>>>(
        in, n_words, W, make_fastdiv((u32)W), make_fastdiv((u32)Y), (u32)Y, (int)X, invert_in, P, out);
    CFE_CHECK_LAUNCH("k_disk_dilate_bits");
    return 0;
}

// fill_holes: flag[root] = 0 for every run that touches the border of its (Y, X) slice (first / last row, first /
// last column).  bits / pref / parent as for the labelling kernels (parent flattened), rows = Z * Y.
__global__ void __launch_bounds__(256)
k_border_flags(const u32* __restrict__ bits, const u32* __restrict__ pref, i64 n_words, int W, FastDiv dW,
               FastDiv dY, u32 Y, int X, const u32* __restrict__ parent, u32* __restrict__ flag)
{
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_words) return;
    const u32 word = __ldg(bits + i);
    if (!word) return;
    const u32 row = fd_div((u32)i, dW);
    const int w = (int)((u32)i - row * (u32)W);
    const u32 z = fd_div(row, dY);
    const u32 y = row - z * Y;
    u32 sel;                                          // bits whose runs touch the border
    if (y == 0 || y == Y - 1) sel = word;
    else {
        sel = 0;
        if (w == 0) sel |= word & 1u;
        if (w == W - 1) sel |= word & (1u << ((X - 1) & 31));
    }
    if (!sel) return;
    const u32 carry = w ? (__ldg(bits + i - 1) >> 31) : 0u;
    const u32 pf = __ldg(pref + i);
    const u32 starts = word & ~((word << 1) | carry);
    while (sel) {
        const int b = __ffs((int)sel) - 1;
        // run covering bit b: the last run start at or before b (or the run carried in from the previous word)
        const u32 id = pf + (u32)__popc(starts & ((b >= 31) ? 0xffffffffu : ((2u << b) - 1u))) - 1u;
        flag[__ldg(parent + id)] = 0u;
        // skip the rest of this run
        const u32 shifted = word >> b;
        const int len = (~shifted == 0u) ? 32 - b : __ffs((int)~shifted) - 1;
        sel &= (b + len >= 32) ? 0u : ~(((1u << len) - 1u) << b);
    }
}

extern "C" int cfe_border_flags(const uint32_t* bits, const uint32_t* pref, int64_t Z, int64_t Y, int64_t X,
                                const uint32_t* parent, uint32_t* flag, cudaStream_t stream)
{
    if (Z <= 0 || Y <= 0 || X <= 0) return 0;
    const int W = (int)((X + 31) / 32);
    const i64 n_words = Z * Y * W;
    if (n_words >= (1LL << 32)) { cfe_set_error("cfe_border_flags: volume too large"); return -1; }
    k_border_flags<<<(unsigned)((n_words + 255) / 256), 256, 0, stream>>>(
        bits, pref, n_words, W, make_fastdiv((u32)W), make_fastdiv((u32)Y), (u32)Y, (int)X, parent, flag);
    CFE_CHECK_LAUNCH("k_border_flags");
    return 0;
}

// ------------------------------------------------------------------------------------------
// 3-D closing with a cube (periosteummask, preprocess.py:408: morphology.binary_closing(perimask,
// morphology.cube(closevoxels))).  A cube is separable: the dilation (erosion) by ones(w, w, w) is three line
// dilations (erosions), one per axis.  scipy centres an even-sized structure at w / 2, which makes the two
// asymmetric: with L = w / 2 and R = w - 1 - L
//     dilation  out[q] = OR  of in[q + e],  e = -R .. +L     (0 outside the image)
//     erosion   out[q] = AND of in[q + e],  e = -L .. +R     (border value 1)  =  ~ dilation'(~in), e = -L .. +R
// k_line_or_bits does one such pass on packed bits: out[q] = OR_{e = -a .. b} src[q + e] along `axis`
// (0 = slices, 1 = rows, 2 = columns), src = in or (invert_in) ~in restricted to the valid bits; invert_out
// complements the result (last pass of the erosion).
__global__ void __launch_bounds__(256)
k_line_or_bits(const u32* __restrict__ in, i64 n_words, int W, FastDiv dW, FastDiv dY, u32 Y, u32 Z, int X, int axis,
               int a, int b, int invert_in, int invert_out, u32* __restrict__ out)
{
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_words) return;
    const u32 row = fd_div((u32)i, dW);
    const int w = (int)((u32)i - row * (u32)W);
    const u32 z = fd_div(row, dY);
    const int y = (int)(row - z * Y);
    const u32 tail = (X & 31) ? ((1u << (X & 31)) - 1u) : 0xffffffffu;      // valid bits of the last word
    auto load = [&](int zz, int yy, int ww) -> u32 {
        if (zz < 0 || zz >= (int)Z || yy < 0 || yy >= (int)Y || ww < 0 || ww >= W) return 0u;
        u32 v = __ldg(in + ((i64)zz * Y + yy) * W + ww);
        if (invert_in) v = ~v;
        if (ww == W - 1) v &= tail;
        return v;
    };
    u32 acc = 0;
    if (axis == 2) {
        // columns: source bit q + e -> shift right by e (e > 0) or left by -e; |e| <= 32: one neighbour word
        const u32 Lw = load((int)z, y, w - 1), M = load((int)z, y, w), Rw = load((int)z, y, w + 1);
        acc = M;
        for (int e = 1; e <= b; ++e) acc |= (e < 32) ? __funnelshift_r(M, Rw, e) : Rw;
        for (int e = 1; e <= a; ++e) acc |= (e < 32) ? __funnelshift_l(Lw, M, e) : Lw;
    } else if (axis == 1) {
        for (int e = -a; e <= b; ++e) acc |= load((int)z, y + e, w);
    } else {
        for (int e = -a; e <= b; ++e) acc |= load((int)z + e, y, w);
    }
    if (invert_out) acc = ~acc;
    if (w == W - 1) acc &= tail;
    out[i] = acc;
}

// One pass: out = OR over the source offsets -a .. +b along `axis` (0 <= a, b <= 32); in != out.
extern "C" int cfe_line_or_bits(const uint32_t* in, int64_t Z, int64_t Y, int64_t X, int axis, int a, int b,
                                int invert_in, int invert_out, uint32_t* out, cudaStream_t stream)
{
    if (Z <= 0 || Y <= 0 || X <= 0) return 0;
    if (axis < 0 || axis > 2 || a < 0 || b < 0 || a > 32 || b > 32 || in == out) {
        cfe_set_error("cfe_line_or_bits: bad arguments");
        return -1;
    }
    const int W = (int)((X + 31) / 32);
    const i64 n_words = Z * Y * W;
    if (n_words >= (1LL << 32)) { cfe_set_error("cfe_line_or_bits: volume too large"); return -1; }
    k_line_or_bits<<<(unsigned)((n_words + 255) / 256), 256, 0, stream>>>(
        in, n_words, W, make_fastdiv((u32)W), make_fastdiv((u32)Y), (u32)Y, (u32)Z, (int)X, axis, a, b, invert_in,
        invert_out, out);
    CFE_CHECK_LAUNCH("k_line_or_bits");
    return 0;
}

// bytes from bits: out[v] = bit v of the packed volume (0 / 1), the inverse of cfe_pack_bits
__global__ void __launch_bounds__(256)
k_unpack_bits(const u32* __restrict__ bits, i64 rows, int X, int W, uint8_t* __restrict__ out)
{
    const i64 g = (i64)blockIdx.x * blockDim.x + threadIdx.x;       // one byte of bits = 8 voxels
    if (g >= rows * W * 4) return;
    const i64 wi = g >> 2;
    const i64 row = wi / W;
    const int c0 = (int)(wi - row * W) * 32 + 8 * (int)(g & 3);
    if (c0 >= X) return;
    const u32 byte = (__ldg(bits + wi) >> (8 * (g & 3))) & 0xffu;
    uint8_t* p = out + row * (i64)X + c0;
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (c0 + k < X) p[k] = (uint8_t)((byte >> k) & 1u);
}

extern "C" int cfe_unpack_bits(const uint32_t* bits, int64_t rows, int64_t X, uint8_t* out, cudaStream_t stream)
{
    if (rows <= 0 || X <= 0) return 0;
    const int W = (int)((X + 31) / 32);
    const i64 n = rows * W * 4;
    k_unpack_bits<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(bits, rows, (int)X, W, out);
    CFE_CHECK_LAUNCH("k_unpack_bits");
    return 0;
}
