// HBM-bound pointwise / resampling / loss / optimizer kernels (NHWC fp32).
// One thread handles 4 consecutive channels of one pixel (float4 when the
// strides allow), grids are sized in multiples of the 148 SMs.
#include "mi_common.cuh"

#include <algorithm>

namespace {

constexpr int TPB = 256;
inline int grid_for(long long work) {
    long long b = (work + TPB - 1) / TPB;
    const long long cap = 148LL * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

#define GRID_STRIDE(i, total) \
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (total); i += (long long)gridDim.x * blockDim.x)

// All NHWC kernels below work on groups of 4 channels: one thread = (pixel, channel group).  With 16-byte
// aligned rows (`vec`) a group is one float4 access (pad lanes of a 4-padded row are never read as data, so
// touching them is harmless); otherwise the group is handled with guarded scalars.
struct F4 { float v[4]; };
__device__ __forceinline__ F4 ld4(const float* p, int valid, bool vec) {
    F4 r;
    if (vec) {
        const float4 t = *reinterpret_cast<const float4*>(p);
        r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) r.v[q] = q < valid ? p[q] : 0.f;
    }
    return r;
}
__device__ __forceinline__ void st4(float* p, const F4& r, int valid, bool vec) {
    // A ragged last group is stored lane by lane: whether the lanes past the channel count are padding or the first
    // channels of the NEXT slice of a concat buffer cannot be told from (pointer, stride, count) -- a [0:2] slice of a
    // 4-channel tensor looks exactly like a 2-channel tensor padded to 4 -- so they are never written.
    if (vec && valid == 4) {
        *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) if (q < valid) p[q] = r.v[q];
    }
}

// store rounded to the TF32 grid (the value is about to be a tensor-core conv operand, see mi_rn_tf32)
__device__ __forceinline__ void st4(float* p, F4 r, int valid, bool vec, int rnd) {
    if (rnd) {
#pragma unroll
        for (int q = 0; q < 4; ++q) r.v[q] = mi_rn_tf32(r.v[q]);
    }
    st4(p, r, valid, vec);
}
// the launchers pass `vec | rnd << 1` in the kernels' `vec` parameter
#define MI_SPLIT_VEC_RND(vec, rnd) const int rnd = (vec) >> 1; (vec) &= 1

// ----------------------------------------------------------------------------- pooling
template <typename IDX>
__global__ void avgpool2_fwd_kernel(const float* __restrict__ x, int ldx, float* __restrict__ y, int ldy, int n, int h,
                                    int wd, int c, int vec) {
    MI_SPLIT_VEC_RND(vec, rnd);
    const int oh = h >> 1, ow = wd >> 1, cg = (c + 3) >> 2;
    const long long total = (long long)n * oh * ow * cg;
    GRID_STRIDE(i, total) {
        const int g = (int)((IDX)i % (IDX)cg);
        IDX p = (IDX)i / (IDX)cg;
        const int ox = (int)(p % (IDX)ow); p /= (IDX)ow;
        const int oy = (int)(p % (IDX)oh);
        const int nn = (int)(p / (IDX)oh);
        const int valid = min(4, c - 4 * g);
        const float* s = x + ((long long)(nn * h + 2 * oy) * wd + 2 * ox) * ldx + 4 * g;
        const F4 a = ld4(s, valid, vec), b = ld4(s + ldx, valid, vec);
        const F4 cc = ld4(s + (long long)wd * ldx, valid, vec), d = ld4(s + (long long)wd * ldx + ldx, valid, vec);
        F4 o;
#pragma unroll
        for (int q = 0; q < 4; ++q) o.v[q] = 0.25f * ((a.v[q] + b.v[q]) + (cc.v[q] + d.v[q]));
        st4(y + ((long long)(nn * oh + oy) * ow + ox) * ldy + 4 * g, o, valid, vec, rnd);
    }
}

template <typename IDX>
__global__ void avgpool2_bwd_kernel(const float* __restrict__ dy, int lddy, float* __restrict__ dx, int lddx,
                                    int accumulate, int n, int h, int wd, int c, int vec) {
    const int oh = h >> 1, ow = wd >> 1, cg = (c + 3) >> 2;
    const long long total = (long long)n * h * wd * cg;
    GRID_STRIDE(i, total) {
        const int g = (int)((IDX)i % (IDX)cg);
        IDX p = (IDX)i / (IDX)cg;
        const int xx = (int)(p % (IDX)wd); p /= (IDX)wd;
        const int yy = (int)(p % (IDX)h);
        const int nn = (int)(p / (IDX)h);
        const int valid = min(4, c - 4 * g);
        F4 v = ld4(dy + ((long long)(nn * oh + (yy >> 1)) * ow + (xx >> 1)) * lddy + 4 * g, valid, vec);
        float* d = dx + ((long long)(nn * h + yy) * wd + xx) * lddx + 4 * g;
        if (accumulate) {
            const F4 o = ld4(d, valid, vec);
#pragma unroll
            for (int q = 0; q < 4; ++q) v.v[q] = o.v[q] + 0.25f * v.v[q];
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) v.v[q] *= 0.25f;
        }
        st4(d, v, valid, vec);
    }
}

__global__ void maxpool2_fwd_kernel(const float* __restrict__ x, int ldx, float* __restrict__ y, int ldy, int n, int h,
                                    int wd, int c) {
    const int oh = h >> 1, ow = wd >> 1;
    const long long total = (long long)n * oh * ow * c;
    GRID_STRIDE(i, total) {
        const int ch = (int)(i % c);
        long long p = i / c;
        const int ox = (int)(p % ow); p /= ow;
        const int oy = (int)(p % oh);
        const int nn = (int)(p / oh);
        const float* s = x + ((long long)(nn * h + 2 * oy) * wd + 2 * ox) * ldx + ch;
        float m = s[0];
        m = fmaxf(m, s[ldx]);
        m = fmaxf(m, s[(long long)wd * ldx]);
        m = fmaxf(m, s[(long long)wd * ldx + ldx]);
        y[((long long)(nn * oh + oy) * ow + ox) * ldy + ch] = m;
    }
}

// gradient goes to the FIRST maximal element in row-major window order (ATen max_pool2d semantics)
__global__ void maxpool2_bwd_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ dy, int lddy,
                                    float* __restrict__ dx, int lddx, int accumulate, int n, int h, int wd, int c) {
    const int oh = h >> 1, ow = wd >> 1;
    const long long total = (long long)n * oh * ow * c;
    GRID_STRIDE(i, total) {
        const int ch = (int)(i % c);
        long long p = i / c;
        const int ox = (int)(p % ow); p /= ow;
        const int oy = (int)(p % oh);
        const int nn = (int)(p / oh);
        const long long base = ((long long)(nn * h + 2 * oy) * wd + 2 * ox);
        const float* s = x + base * ldx + ch;
        const float v[4] = {s[0], s[ldx], s[(long long)wd * ldx], s[(long long)wd * ldx + ldx]};
        int arg = 0;
        float m = v[0];
#pragma unroll
        for (int q = 1; q < 4; ++q) if (v[q] > m) { m = v[q]; arg = q; }
        const float g = dy[((long long)(nn * oh + oy) * ow + ox) * lddy + ch];
        float* d = dx + base * lddx + ch;
        const long long off[4] = {0, lddx, (long long)wd * lddx, (long long)wd * lddx + lddx};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float gv = (q == arg) ? g : 0.f;
            d[off[q]] = accumulate ? d[off[q]] + gv : gv;
        }
    }
}

// ----------------------------------------------------------------------------- bilinear x2
// source coordinate of output index o (ATen upsample_bilinear2d, scale factor 2)
// (`scale` = (in - 1) / (2 in - 1) for align_corners, computed ONCE on the host in the same float arithmetic: the
// division used to be evaluated in every call, 2 per output pixel forward and 14 per input pixel backward)
static inline float up2_scale(int in_size) {
    const int out_size = in_size * 2;
    return out_size > 1 ? (float)(in_size - 1) / (float)(out_size - 1) : 0.f;
}
__device__ __forceinline__ void up2_src(int o, int in_size, int align, float scale, int& i0, int& i1, float& t) {
    float s;
    if (align) {
        s = scale * (float)o;
    } else {
        s = 0.5f * ((float)o + 0.5f) - 0.5f;
        if (s < 0.f) s = 0.f;
    }
    i0 = (int)s;
    if (i0 > in_size - 1) i0 = in_size - 1;
    i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    t = s - (float)i0;
}

// Windowed form: the low-resolution buffer x holds rows [ly0, ly0+h) x cols [lx0, lx0+wd) of a full_h x full_w grid and
// the output buffer y holds rows [hy0, hy0+oh) x cols [hx0, hx0+ow) of its x2 upsampling; interpolation weights are
// those of the FULL grid (align_corners=True depends on the full size), so a cropped evaluation reproduces the
// full one bit for bit wherever the sources lie inside the low-resolution window (the host checks that they do).
// The plain x2 upsample is the window that covers everything.
struct UpWin { int full_h, full_w, ly0, lx0, hy0, hx0, oh, ow; float sy, sx; };

template <typename IDX>
__global__ void upsample2_fwd_kernel(const float* __restrict__ x, int ldx, float* __restrict__ y, int ldy, int n, int h,
                                     int wd, int c, int align, int vec, UpWin g) {
    MI_SPLIT_VEC_RND(vec, rnd);
    const int oh = g.oh, ow = g.ow, cg = (c + 3) >> 2;
    const long long total = (long long)n * oh * ow * cg;
    GRID_STRIDE(i, total) {
        const int gi = (int)((IDX)i % (IDX)cg);
        IDX p = (IDX)i / (IDX)cg;
        const int ox = (int)(p % (IDX)ow); p /= (IDX)ow;
        const int oy = (int)(p % (IDX)oh);
        const int nn = (int)(p / (IDX)oh);
        const int valid = min(4, c - 4 * gi);
        int y0, y1, x0, x1; float ty, tx;
        up2_src(oy + g.hy0, g.full_h, align, g.sy, y0, y1, ty);
        up2_src(ox + g.hx0, g.full_w, align, g.sx, x0, x1, tx);
        y0 = min(max(y0 - g.ly0, 0), h - 1); y1 = min(max(y1 - g.ly0, 0), h - 1);
        x0 = min(max(x0 - g.lx0, 0), wd - 1); x1 = min(max(x1 - g.lx0, 0), wd - 1);
        const float* b = x + (long long)nn * h * wd * ldx + 4 * gi;
        const F4 v00 = ld4(b + ((long long)y0 * wd + x0) * ldx, valid, vec), v01 = ld4(b + ((long long)y0 * wd + x1) * ldx, valid, vec);
        const F4 v10 = ld4(b + ((long long)y1 * wd + x0) * ldx, valid, vec), v11 = ld4(b + ((long long)y1 * wd + x1) * ldx, valid, vec);
        F4 o;
#pragma unroll
        for (int q = 0; q < 4; ++q)
            o.v[q] = (1.f - ty) * ((1.f - tx) * v00.v[q] + tx * v01.v[q]) + ty * ((1.f - tx) * v10.v[q] + tx * v11.v[q]);
        st4(y + ((long long)(nn * oh + oy) * ow + ox) * ldy + 4 * gi, o, valid, vec, rnd);
    }
}

// gather form of the transpose: each input pixel collects from the output pixels that read it.  Output o reads
// inputs (i0,i1); the outputs that can touch input i lie within [2i-3, 2i+3]; membership is tested exactly.
// Optional mask: `mask_y` is the post-activation tensor whose x2 upsampling is being differentiated (same shape as
// dx); the result is then the gradient w.r.t. its PRE-activation, which saves the separate act_bwd pass.
template <typename IDX>
__global__ void upsample2_bwd_kernel(const float* __restrict__ dy, int lddy, float* __restrict__ dx, int lddx,
                                     int accumulate, int n, int h, int wd, int c, int align, int vec, UpWin g,
                                     const float* __restrict__ mask_y, int ldmask, int mask_act, float mask_slope) {
    MI_SPLIT_VEC_RND(vec, rnd);
    const int oh = g.oh, ow = g.ow, cg = (c + 3) >> 2;
    const long long total = (long long)n * h * wd * cg;
    GRID_STRIDE(i, total) {
        const int gi = (int)((IDX)i % (IDX)cg);
        IDX p = (IDX)i / (IDX)cg;
        const int xx = (int)(p % (IDX)wd); p /= (IDX)wd;
        const int yy = (int)(p % (IDX)h);
        const int nn = (int)(p / (IDX)h);
        const int valid = min(4, c - 4 * gi);
        const int gy = yy + g.ly0, gx = xx + g.lx0;          // position on the full low-resolution grid
        float wy[6]; int oy_[6]; int ny = 0;
        for (int o = max(g.hy0, 2 * gy - 3); o <= min(g.hy0 + oh - 1, 2 * gy + 3); ++o) {
            int a, b; float t; up2_src(o, g.full_h, align, g.sy, a, b, t);
            float wgt = 0.f;
            if (a == gy) wgt += 1.f - t;
            if (b == gy) wgt += t;
            if (wgt != 0.f && ny < 6) { wy[ny] = wgt; oy_[ny] = o - g.hy0; ++ny; }
        }
        float wx[6]; int ox_[6]; int nx = 0;
        for (int o = max(g.hx0, 2 * gx - 3); o <= min(g.hx0 + ow - 1, 2 * gx + 3); ++o) {
            int a, b; float t; up2_src(o, g.full_w, align, g.sx, a, b, t);
            float wgt = 0.f;
            if (a == gx) wgt += 1.f - t;
            if (b == gx) wgt += t;
            if (wgt != 0.f && nx < 6) { wx[nx] = wgt; ox_[nx] = o - g.hx0; ++nx; }
        }
        F4 acc;
#pragma unroll
        for (int q = 0; q < 4; ++q) acc.v[q] = 0.f;
        for (int a = 0; a < ny; ++a)
            for (int b = 0; b < nx; ++b) {
                const F4 v = ld4(dy + ((long long)(nn * oh + oy_[a]) * ow + ox_[b]) * lddy + 4 * gi, valid, vec);
                const float wgt = wy[a] * wx[b];
#pragma unroll
                for (int q = 0; q < 4; ++q) acc.v[q] += wgt * v.v[q];
            }
        float* d = dx + ((long long)(nn * h + yy) * wd + xx) * lddx + 4 * gi;
        if (accumulate) {
            const F4 o = ld4(d, valid, vec);
#pragma unroll
            for (int q = 0; q < 4; ++q) acc.v[q] += o.v[q];
        }
        if (mask_y) {
            const float* mp = mask_y + ((long long)(nn * h + yy) * wd + xx) * ldmask + 4 * gi;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (q < valid) acc.v[q] *= mi_act_grad(mp[q], mask_act, mask_slope);
        }
        st4(d, acc, valid, vec, rnd);
    }
}

// ---- strip forms of the two kernels above (default; MI_B200_UPSAMPLE_STRIP=0 selects the per-pixel forms) ----------
// One thread = one column position (x, 4-channel group) walking a STRIP of rows; blockIdx.y = strip, blockIdx.z = image.
// What was per output value in the per-pixel form is now per thread (the index decomposition: one division instead of
// six; the horizontal source coordinates / weights) or per block (the vertical ones: a table in shared memory), and a
// source row is fetched once per thread for all the rows of the strip it feeds.  The per-pixel forms were issue-bound
// at 0.4 (forward) and 0.3 (backward) of the HBM rate (28 / 40 us for 75 MB on the 137x233 -> 258x450 Subnet upsample);
// ncu still counted 130 (forward) and 500 (backward) instructions per 16 bytes stored in the first strip version, which
// recomputed the vertical weights per thread and row and carried 64-bit offsets -- hence the table and the 32-bit
// offsets inside an image (the launcher checks that they fit).
// Rows per thread (`strip`): UP_STRIP_MAX at most, fewer when the launch would otherwise not fill the chip
// (up_strip_grid).  VEC / RND are the two bits of the per-pixel kernels' `vec` parameter, fixed at compile time here.
constexpr int UP_STRIP_MAX = 8;
constexpr int UP_SLOTS = UP_STRIP_MAX / 2 + 2;          // source rows a strip of output rows can touch
constexpr int UP_TPB = 256;

// forward.  The strip's source rows are blended horizontally ONCE, hx(row) = (1-tx) x[row][x0] + tx x[row][x1], all
// loads of the strip in flight together, and parked in the thread's own column of a shared-memory table (a register
// file that can be indexed); an output row then is (1-ty) hx[slot0] + ty hx[slot1] with the block's row table.
template <bool VEC, bool RND>
__global__ void __launch_bounds__(UP_TPB)
upsample2_fwd_strip_kernel(const float* __restrict__ x, int ldx, float* __restrict__ y, int ldy, int h, int wd, int c,
                           int align, UpWin g, int strip) {
    __shared__ float4 s_h[UP_SLOTS][UP_TPB];
    __shared__ int s_slot0[UP_STRIP_MAX], s_slot1[UP_STRIP_MAX];
    __shared__ float s_ty[UP_STRIP_MAX];
    __shared__ int s_base, s_nrows;
    const int oy0 = blockIdx.y * strip, rows = min(strip, g.oh - oy0);
    if (threadIdx.x == 0) {
        int first = 0, last = 0;
        for (int r = 0; r < rows; ++r) {
            int y0, y1; float ty;
            up2_src(oy0 + r + g.hy0, g.full_h, align, g.sy, y0, y1, ty);
            y0 = min(max(y0 - g.ly0, 0), h - 1); y1 = min(max(y1 - g.ly0, 0), h - 1);
            if (r == 0) first = y0;
            s_slot0[r] = min(y0 - first, UP_SLOTS - 1); s_slot1[r] = min(y1 - first, UP_SLOTS - 1); s_ty[r] = ty;
            last = y1;
        }
        s_base = first; s_nrows = min(last - first + 1, UP_SLOTS);
    }
    __syncthreads();
    const int cg = (c + 3) >> 2;
    const unsigned j = blockIdx.x * UP_TPB + threadIdx.x;
    const int ox = (int)(j / (unsigned)cg), gi = (int)(j - (unsigned)ox * (unsigned)cg);
    if (ox >= g.ow) return;
    const int valid = min(4, c - 4 * gi);
    int x0, x1; float tx;
    up2_src(ox + g.hx0, g.full_w, align, g.sx, x0, x1, tx);
    x0 = min(max(x0 - g.lx0, 0), wd - 1); x1 = min(max(x1 - g.lx0, 0), wd - 1);
    const float* px = x + (size_t)blockIdx.z * h * wd * ldx + 4 * gi;
    const int base = s_base, nrows = s_nrows, rowpitch = wd * ldx;
    const int off0 = base * rowpitch + x0 * ldx, off1 = base * rowpitch + x1 * ldx;
    F4 v0[UP_SLOTS], v1[UP_SLOTS];
#pragma unroll
    for (int q = 0; q < UP_SLOTS; ++q)
        if (q < nrows) { v0[q] = ld4(px + off0 + q * rowpitch, valid, VEC); v1[q] = ld4(px + off1 + q * rowpitch, valid, VEC); }
#pragma unroll
    for (int q = 0; q < UP_SLOTS; ++q)
        if (q < nrows)
            s_h[q][threadIdx.x] = make_float4((1.f - tx) * v0[q].v[0] + tx * v1[q].v[0], (1.f - tx) * v0[q].v[1] + tx * v1[q].v[1],
                                              (1.f - tx) * v0[q].v[2] + tx * v1[q].v[2], (1.f - tx) * v0[q].v[3] + tx * v1[q].v[3]);
    float* po = y + (size_t)blockIdx.z * g.oh * g.ow * ldy + (oy0 * g.ow + ox) * ldy + 4 * gi;
    const int opitch = g.ow * ldy;
    for (int r = 0; r < rows; ++r) {
        const float4 n0 = s_h[s_slot0[r]][threadIdx.x], n1 = s_h[s_slot1[r]][threadIdx.x];
        const float ty = s_ty[r];
        F4 o;
        o.v[0] = (1.f - ty) * n0.x + ty * n1.x; o.v[1] = (1.f - ty) * n0.y + ty * n1.y;
        o.v[2] = (1.f - ty) * n0.z + ty * n1.z; o.v[3] = (1.f - ty) * n0.w + ty * n1.w;
        st4(po + r * opitch, o, valid, VEC, RND);
    }
}

// backward: `strip` input rows per thread.  The thread walks the output rows that feed its strip in ascending order;
// per output row it reduces the horizontal contributions r = sum_b wx[b] dy[o][b] once and adds (1-t) r / t r to the two
// input rows the output row was interpolated from, which only ever are the current row and the next one: two running
// accumulators, a row is finished (accumulate / activation mask / store) when the walk leaves it.  Every dy row is read
// once per thread instead of once per input row it feeds.
constexpr int UP_CAND = 6;                               // candidate output columns 2 gx - 2 .. 2 gx + 3 of input column gx
constexpr int UP_WALK = 2 * UP_STRIP_MAX + UP_CAND;      // output rows that can feed a strip
template <bool VEC, bool RND>
__global__ void __launch_bounds__(UP_TPB)
upsample2_bwd_strip_kernel(const float* __restrict__ dy, int lddy, float* __restrict__ dx, int lddx, int accumulate,
                           int h, int wd, int c, int align, UpWin g, const float* __restrict__ mask_y, int ldmask,
                           int mask_act, float mask_slope, int strip) {
    __shared__ int s_a[UP_WALK], s_b[UP_WALK];
    __shared__ float s_t[UP_WALK];
    const int oh = g.oh, ow = g.ow;
    const int yy0 = blockIdx.y * strip, yy1 = min(yy0 + strip, h) - 1;        // local rows of the strip, inclusive
    const int gy0 = yy0 + g.ly0, gy1 = yy1 + g.ly0;                           // ... on the full low-resolution grid
    // The source coordinate of output o lies in [o / 2 - 1 / 2, o / 2] for either convention, so outputs below 2 gy - 2
    // interpolate from rows gy - 2 and gy - 1 at most and never feed gy; 2 gy + 3 can, when the product rounds below
    // gy + 1 (the same holds for the columns).
    const int o_lo = max(g.hy0, 2 * gy0 - 2), o_hi = min(g.hy0 + oh - 1, 2 * gy1 + 3);
    if ((int)threadIdx.x <= o_hi - o_lo && threadIdx.x < UP_WALK) {
        int a, b; float t; up2_src(o_lo + threadIdx.x, g.full_h, align, g.sy, a, b, t);
        s_a[threadIdx.x] = a; s_b[threadIdx.x] = b; s_t[threadIdx.x] = t;
    }
    __syncthreads();
    const int cg = (c + 3) >> 2;
    const unsigned j = blockIdx.x * UP_TPB + threadIdx.x;
    const int xx = (int)(j / (unsigned)cg), gi = (int)(j - (unsigned)xx * (unsigned)cg);
    if (xx >= wd) return;
    const int valid = min(4, c - 4 * gi);
    // horizontal weights of the candidate output columns (zero: not a contributor).  Every candidate is LOADED, from a
    // clamped address when it cannot contribute, and enters with weight zero: loads under a per-thread condition
    // became one divergent branch per column and row, each waiting for its own load.
    const int gx = xx + g.lx0, oxb = 2 * gx - 2;
    float wx[UP_CAND];
    int col[UP_CAND];
#pragma unroll
    for (int k = 0; k < UP_CAND; ++k) {
        const int o = oxb + k;
        int a, b; float t; up2_src(o, g.full_w, align, g.sx, a, b, t);
        float wgt = 0.f;
        if (a == gx) wgt += 1.f - t;
        if (b == gx) wgt += t;
        wx[k] = (o >= g.hx0 && o < g.hx0 + ow) ? wgt : 0.f;
        col[k] = min(max(o - g.hx0, 0), ow - 1) * lddy;
    }
    const float* pin = dy + (size_t)blockIdx.z * oh * ow * lddy + 4 * gi;
    float* pdx = dx + (size_t)blockIdx.z * h * wd * lddx + xx * lddx + 4 * gi;
    const float* pm = mask_y ? mask_y + (size_t)blockIdx.z * h * wd * ldmask + xx * ldmask + 4 * gi : nullptr;
    const int ipitch = ow * lddy, dpitch = wd * lddx, mpitch = wd * ldmask;
    int cur = gy0;                                                   // acc0 belongs to row cur, acc1 to row cur + 1
    F4 acc0, acc1;
#pragma unroll
    for (int q = 0; q < 4; ++q) acc0.v[q] = acc1.v[q] = 0.f;
    auto finish = [&](int grow, F4 acc) {
        const int yy = grow - g.ly0;
        float* d = pdx + yy * dpitch;
        if (accumulate) {
            const F4 o = ld4(d, valid, VEC);
#pragma unroll
            for (int q = 0; q < 4; ++q) acc.v[q] += o.v[q];
        }
        if (pm) {
            const F4 m = ld4(pm + yy * mpitch, valid, VEC);
#pragma unroll
            for (int q = 0; q < 4; ++q) acc.v[q] *= mi_act_grad(m.v[q], mask_act, mask_slope);
        }
        st4(d, acc, valid, VEC, RND);
    };
    for (int o = o_lo; o <= o_hi; ++o) {
        const int a = s_a[o - o_lo], b = s_b[o - o_lo];
        if (b < gy0 || a > gy1) continue;                            // feeds rows outside the strip only
        const float t = s_t[o - o_lo];
        while (cur < a) {                                            // (uniform over the block)
            finish(cur, acc0);
            acc0 = acc1;
#pragma unroll
            for (int q = 0; q < 4; ++q) acc1.v[q] = 0.f;
            ++cur;
        }
        const float* prow = pin + (o - g.hy0) * ipitch;
        F4 v[UP_CAND];
#pragma unroll
        for (int k = 0; k < UP_CAND; ++k) v[k] = ld4(prow + col[k], valid, VEC);
        F4 r;
#pragma unroll
        for (int q = 0; q < 4; ++q) r.v[q] = wx[0] * v[0].v[q];
#pragma unroll
        for (int k = 1; k < UP_CAND; ++k)
#pragma unroll
            for (int q = 0; q < 4; ++q) r.v[q] += wx[k] * v[k].v[q];
        // a >= cur - 1 here: a == cur - 1 only for the rows above the strip (a < gy0 = first cur), whose own share is dropped
        const float w0 = 1.f - t;
        if (a == cur) {
#pragma unroll
            for (int q = 0; q < 4; ++q) acc0.v[q] += w0 * r.v[q];
            if (b == a) {
#pragma unroll
                for (int q = 0; q < 4; ++q) acc0.v[q] += t * r.v[q];
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) acc1.v[q] += t * r.v[q];
            }
        } else if (b == cur) {                                       // a == cur - 1: above the strip
#pragma unroll
            for (int q = 0; q < 4; ++q) acc0.v[q] += t * r.v[q];
        }
    }
    while (cur <= gy1) {
        finish(cur, acc0);
        acc0 = acc1;
#pragma unroll
        for (int q = 0; q < 4; ++q) acc1.v[q] = 0.f;
        ++cur;
    }
}

// dst window (+)= src window, both NHWC buffers with their own extents (crop of a region of interest and its adjoint)
template <typename IDX>
__global__ void window_copy_kernel(const float* __restrict__ s, int lds, int sh, int sw, int sy0, int sx0,
                                   float* __restrict__ d, int ldd, int dh, int dw, int dy0, int dx0, int n, int h,
                                   int w, int c, int accumulate, int vec) {
    const int cg = (c + 3) >> 2;
    const long long total = (long long)n * h * w * cg;
    GRID_STRIDE(i, total) {
        const int gi = (int)((IDX)i % (IDX)cg);
        IDX p = (IDX)i / (IDX)cg;
        const int xx = (int)(p % (IDX)w); p /= (IDX)w;
        const int yy = (int)(p % (IDX)h);
        const int nn = (int)(p / (IDX)h);
        const int valid = min(4, c - 4 * gi);
        F4 v = ld4(s + (((long long)nn * sh + sy0 + yy) * sw + sx0 + xx) * lds + 4 * gi, valid, vec);
        float* q4 = d + (((long long)nn * dh + dy0 + yy) * dw + dx0 + xx) * ldd + 4 * gi;
        if (accumulate) {
            const F4 o = ld4(q4, valid, vec);
#pragma unroll
            for (int q = 0; q < 4; ++q) v.v[q] += o.v[q];
        }
        st4(q4, v, valid, vec);
    }
}

// ----------------------------------------------------------------------------- simple pointwise
template <typename IDX>
__global__ void add_kernel(const float* __restrict__ a, int lda, const float* __restrict__ b, int ldb,
                           float* __restrict__ y, int ldy, long long pixels, int c, int vec) {
    MI_SPLIT_VEC_RND(vec, rnd);
    const int cg = (c + 3) >> 2;
    const long long total = pixels * cg;
    GRID_STRIDE(i, total) {
        const int g = (int)((IDX)i % (IDX)cg);
        const IDX p = (IDX)i / (IDX)cg;
        const int valid = min(4, c - 4 * g);
        const F4 u = ld4(a + (long long)p * lda + 4 * g, valid, vec), v = ld4(b + (long long)p * ldb + 4 * g, valid, vec);
        F4 o;
#pragma unroll
        for (int q = 0; q < 4; ++q) o.v[q] = u.v[q] + v.v[q];
        st4(y + (long long)p * ldy + 4 * g, o, valid, vec, rnd);
    }
}

template <typename IDX>
__global__ void copy_kernel(const float* __restrict__ s, int lds, float* __restrict__ d, int ldd, int accumulate,
                            long long pixels, int c, int vec) {
    const int cg = (c + 3) >> 2;
    const long long total = pixels * cg;
    GRID_STRIDE(i, total) {
        const int g = (int)((IDX)i % (IDX)cg);
        const IDX p = (IDX)i / (IDX)cg;
        const int valid = min(4, c - 4 * g);
        F4 v = ld4(s + (long long)p * lds + 4 * g, valid, vec);
        float* q4 = d + (long long)p * ldd + 4 * g;
        if (accumulate) {
            const F4 o = ld4(q4, valid, vec);
#pragma unroll
            for (int q = 0; q < 4; ++q) v.v[q] += o.v[q];
        }
        st4(q4, v, valid, vec);
    }
}

template <typename IDX>
__global__ void act_bwd_kernel(float* __restrict__ dy, int lddy, const float* __restrict__ y, int ldy, int act,
                               float slope, long long pixels, int c, int vec, int rnd) {
    const int cg = (c + 3) >> 2;
    const long long total = pixels * cg;
    GRID_STRIDE(i, total) {
        const int g = (int)((IDX)i % (IDX)cg);
        const IDX p = (IDX)i / (IDX)cg;
        const int valid = min(4, c - 4 * g);
        F4 d = ld4(dy + (long long)p * lddy + 4 * g, valid, vec);
        const F4 v = ld4(y + (long long)p * ldy + 4 * g, valid, vec);
#pragma unroll
        for (int q = 0; q < 4; ++q) d.v[q] *= mi_act_grad(v.v[q], act, slope);
        if (rnd) {
#pragma unroll
            for (int q = 0; q < 4; ++q) d.v[q] = mi_rn_tf32(d.v[q]);
        }
        st4(dy + (long long)p * lddy + 4 * g, d, valid, vec);
    }
}

__global__ void fill_kernel(float* __restrict__ p, float v, long long count) {
    GRID_STRIDE(i, count) p[i] = v;
}

__global__ void axpby_kernel(const float* __restrict__ x, float a, float* __restrict__ y, float b, long long count) {
    GRID_STRIDE(i, count) y[i] = a * x[i] + (b == 0.f ? 0.f : b * y[i]);
}

__global__ void addcmul_kernel(float* __restrict__ y, float a, const float* __restrict__ x1,
                               const float* __restrict__ x2, long long count) {
    GRID_STRIDE(i, count) y[i] = fmaf(a * x1[i], x2[i], y[i]);
}

// ----------------------------------------------------------------------------- frames <-> NHWC
__device__ __forceinline__ int reflect_idx(int i, int nsz) {
    if (nsz == 1) return 0;
    while (i < 0 || i >= nsz) {
        if (i < 0) i = -i;
        if (i >= nsz) i = 2 * (nsz - 1) - i;
    }
    return i;
}

__global__ void frames_to_canvas_kernel(const float* __restrict__ f0, const float* __restrict__ f1,
                                        float* __restrict__ canvas, int ldc, int n, int h, int wd, int ch, int cw,
                                        int pad_top, int pad_left, int mode, int rnd) {
    const long long total = (long long)n * ch * cw;
    GRID_STRIDE(i, total) {
        long long p = i;
        const int x = (int)(p % cw); p /= cw;
        const int y = (int)(p % ch);
        const int nn = (int)(p / ch);
        int sy = y - pad_top, sx = x - pad_left;
        if (mode == 0) {
            sy = min(max(sy, 0), h - 1);
            sx = min(max(sx, 0), wd - 1);
        } else {
            sy = reflect_idx(sy, h);
            sx = reflect_idx(sx, wd);
        }
        const long long plane = (long long)h * wd;
        const long long src = (long long)nn * 3 * plane + (long long)sy * wd + sx;
        float* d = canvas + i * ldc;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float a = f0[src + c * plane], b = f1[src + c * plane];
            d[c] = rnd ? mi_rn_tf32(a) : a;
            d[3 + c] = rnd ? mi_rn_tf32(b) : b;
        }
        for (int c = 6; c < ldc; ++c) d[c] = 0.f;
    }
}

__global__ void nhwc_window_to_nchw_kernel(const float* __restrict__ s, int lds, float* __restrict__ d, int n, int hs,
                                           int ws, int y0, int x0, int h, int wd, int c) {
    const long long total = (long long)n * c * h * wd;
    GRID_STRIDE(i, total) {
        long long p = i;
        const int x = (int)(p % wd); p /= wd;
        const int y = (int)(p % h); p /= h;
        const int cc = (int)(p % c);
        const int nn = (int)(p / c);
        d[i] = s[((long long)(nn * hs + y0 + y) * ws + x0 + x) * lds + cc];
    }
}

__global__ void nchw_to_nhwc_window_kernel(const float* __restrict__ s, float* __restrict__ d, int ldd, int n, int hs,
                                           int ws, int y0, int x0, int h, int wd, int c) {
    const long long total = (long long)n * c * h * wd;
    GRID_STRIDE(i, total) {
        long long p = i;
        const int x = (int)(p % wd); p /= wd;
        const int y = (int)(p % h); p /= h;
        const int cc = (int)(p % c);
        const int nn = (int)(p / c);
        d[((long long)(nn * hs + y0 + y) * ws + x0 + x) * ldd + cc] = s[i];
    }
}

// ----------------------------------------------------------------------------- loss / metrics
__global__ void loss_fwd_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                    float* __restrict__ grad, float* __restrict__ loss_out, long long count, int kind,
                                    float weight) {
    const float inv = 1.f / (float)count;
    float local = 0.f;
    GRID_STRIDE(i, count) {
        const float d = pred[i] - target[i];
        if (kind == 0) {
            local += fabsf(d);
            if (grad) grad[i] = weight * inv * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
        } else {
            local += d * d;
            if (grad) grad[i] = weight * inv * 2.f * d;
        }
    }
    __shared__ float red[TPB / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < TPB / 32 ? red[threadIdx.x] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) atomicAdd(loss_out, weight * inv * v);
    }
}

__global__ void psnr_accumulate_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                       double* __restrict__ sq_out, long long count) {
    double local = 0.0;
    GRID_STRIDE(i, count) {
        const float qp = rintf(fminf(fmaxf(pred[i] * 255.f, 0.f), 255.f));
        const float qt = rintf(fminf(fmaxf(target[i] * 255.f, 0.f), 255.f));
        const float d = (qp - qt) / 255.f;
        local += (double)(d * d);
    }
    __shared__ double red[TPB / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        double v = 0.0;
        for (int q = 0; q < TPB / 32; ++q) v += red[q];
        atomicAdd(sq_out, v);
    }
}

// ----------------------------------------------------------------------------- SSIM
// utils.py:195-204 -> pytorch_msssim/__init__.py:19-75 (size_average, val_range=255): both images are quantised to
// 8 bits, the five 11x11 Gaussian-windowed moments (x, y, xx, yy, xy) are "valid" convolutions, and the SSIM map is
// averaged over every channel and pixel.  One CTA owns a 32x16 tile of one channel plane: the quantised
// (32+10)x(16+10) inputs are staged in shared memory once, the window is applied separably (rows, then columns), and
// the tile's SSIM sum goes to one double atomic.  All inputs are integers <= 255, so the products are exact in fp32.
constexpr int SS_TX = 32, SS_TY = 16, SS_MAXWIN = 11;
struct SsimWindow { float g[SS_MAXWIN]; };

__global__ void __launch_bounds__(SS_TX* SS_TY)
ssim_accumulate_kernel(const float* __restrict__ pred, const float* __restrict__ target, double* __restrict__ sum_out,
                       int h, int w, int win, SsimWindow wnd, float c1, float c2) {
    __shared__ float sp[SS_TY + SS_MAXWIN - 1][SS_TX + SS_MAXWIN - 1];
    __shared__ float st[SS_TY + SS_MAXWIN - 1][SS_TX + SS_MAXWIN - 1];
    __shared__ float hz[5][SS_TY + SS_MAXWIN - 1][SS_TX];
    __shared__ double red[SS_TX * SS_TY / 32];
    const int tid = threadIdx.x;
    const int ox0 = blockIdx.x * SS_TX, oy0 = blockIdx.y * SS_TY;
    const int oh = h - win + 1, ow = w - win + 1;
    const size_t plane = (size_t)blockIdx.z * h * w;
    const int rows = SS_TY + win - 1, cols = SS_TX + win - 1;
    for (int i = tid; i < rows * cols; i += SS_TX * SS_TY) {
        const int r = i / cols, c = i % cols;
        const int y = oy0 + r, x = ox0 + c;
        float a = 0.f, b = 0.f;
        if (y < h && x < w) {
            a = rintf(fminf(fmaxf(pred[plane + (size_t)y * w + x] * 255.f, 0.f), 255.f));
            b = rintf(fminf(fmaxf(target[plane + (size_t)y * w + x] * 255.f, 0.f), 255.f));
        }
        sp[r][c] = a;
        st[r][c] = b;
    }
    __syncthreads();
    for (int i = tid; i < rows * SS_TX; i += SS_TX * SS_TY) {
        const int r = i / SS_TX, c = i % SS_TX;
        float m1 = 0.f, m2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
        for (int k = 0; k < win; ++k) {
            const float g = wnd.g[k], a = sp[r][c + k], b = st[r][c + k];
            m1 += g * a; m2 += g * b; s11 += g * (a * a); s22 += g * (b * b); s12 += g * (a * b);
        }
        hz[0][r][c] = m1; hz[1][r][c] = m2; hz[2][r][c] = s11; hz[3][r][c] = s22; hz[4][r][c] = s12;
    }
    __syncthreads();
    const int tx = tid % SS_TX, ty = tid / SS_TX;
    double local = 0.0;
    if (oy0 + ty < oh && ox0 + tx < ow) {
        float m1 = 0.f, m2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
        for (int k = 0; k < win; ++k) {
            const float g = wnd.g[k];
            m1 += g * hz[0][ty + k][tx]; m2 += g * hz[1][ty + k][tx]; s11 += g * hz[2][ty + k][tx];
            s22 += g * hz[3][ty + k][tx]; s12 += g * hz[4][ty + k][tx];
        }
        const float mu11 = m1 * m1, mu22 = m2 * m2, mu12 = m1 * m2;
        const float v1 = 2.f * (s12 - mu12) + c2;
        const float v2 = (s11 - mu11) + (s22 - mu22) + c2;
        local = (double)(((2.f * mu12 + c1) * v1) / ((mu11 + mu22 + c1) * v2));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((tid & 31) == 0) red[tid >> 5] = local;
    __syncthreads();
    if (tid == 0) {
        double v = 0.0;
        for (int q = 0; q < SS_TX * SS_TY / 32; ++q) v += red[q];
        atomicAdd(sum_out, v);
    }
}

// ----------------------------------------------------------------------------- inner-loop rules (flat arena)
__global__ void inner_update_kernel(const float* __restrict__ w_in, const float* __restrict__ g,
                                    float* __restrict__ w_out, float* __restrict__ exp_avg,
                                    float* __restrict__ exp_avg_sq, const float* __restrict__ lr, int lr_per_element,
                                    int lr_stride, int num_step, const int32_t* __restrict__ seg,
                                    const uint8_t* __restrict__ skip, long long count, int rule, float bc1,
                                    float bc2_sqrt) {
    const float b1 = 0.9f, b2 = 0.99f, eps = 1e-8f;
    GRID_STRIDE(i, count) {
        const int t = seg[i >> 10];
        const float w = w_in[i];
        if (t < 0 || (skip && skip[t])) { w_out[i] = w; continue; }
        const float l = lr_per_element ? lr[i] : lr[(long long)t * lr_stride + num_step];
        const float gi = g[i];
        float o;
        if (rule == 0) {
            o = w - l * gi;
        } else if (rule == 1) {
            const float m = b1 * exp_avg[i] + (1.f - b1) * gi;
            const float v = b2 * exp_avg_sq[i] + (1.f - b2) * gi * gi;
            exp_avg[i] = m;
            exp_avg_sq[i] = v;
            const float denom = sqrtf(v) / bc2_sqrt + eps;
            o = w - (l / bc1) * m / denom;
        } else if (rule == 2) {
            const float m = b1 * exp_avg[i] + (1.f - b1) * gi;
            exp_avg[i] = m;
            o = w - (l / bc1) * m / (fabsf(gi) + eps);
        } else {
            const float m = (1.f - b1) * gi;
            o = w - (l / bc1) * m / (fabsf(gi) + eps);
        }
        w_out[i] = o;
    }
}

__global__ void outer_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                  float* __restrict__ v, long long count, int kind, float lr, float b1, float b2,
                                  float eps, float wd, float step_size, float bc2_sqrt) {
    GRID_STRIDE(i, count) {
        float gi = g[i];
        float pi = p[i];
        if (wd != 0.f) gi += wd * pi;
        if (kind == 0) {
            pi -= lr * gi;
        } else if (kind == 1) {  // torch.optim.Adam (amsgrad off)
            const float mi = b1 * m[i] + (1.f - b1) * gi;
            const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
            m[i] = mi; v[i] = vi;
            const float denom = sqrtf(vi) / bc2_sqrt + eps;
            pi -= step_size * (mi / denom);
        } else {                 // torch.optim.Adamax
            const float mi = b1 * m[i] + (1.f - b1) * gi;
            const float ui = fmaxf(b2 * v[i], fabsf(gi) + eps);
            m[i] = mi; v[i] = ui;
            pi -= step_size * (mi / ui);
        }
        p[i] = pi;
    }
}

__global__ void segment_dot_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                   const int32_t* __restrict__ seg, float* __restrict__ out, long long count) {
    // one block per 1024-float chunk (chunks never straddle tensors)
    const long long chunk = blockIdx.x;
    const long long base = chunk << 10;
    const int t = seg[chunk];
    if (t < 0) return;
    float local = 0.f;
    for (int q = threadIdx.x; q < 1024; q += blockDim.x) {
        const long long i = base + q;
        if (i < count) local += b ? a[i] * b[i] : a[i];
    }
    __shared__ float red[TPB / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        float v = 0.f;
        for (int q = 0; q < TPB / 32; ++q) v += red[q];
        atomicAdd(out + t, v);
    }
}

// y (+)= alpha * s(t) * x over a flat arena, with s(t) = scale[t] for the tensors selected by `mask` (all when NULL)
// and 1 otherwise; padding chunks are left untouched.  L2F: theta' = gamma (.) theta and dL/dtheta += gamma (.) G.
__global__ void segment_scale_kernel(const float* __restrict__ x, const float* __restrict__ scale,
                                     const int32_t* __restrict__ seg, const float* __restrict__ mask,
                                     float* __restrict__ y, float alpha, int accumulate, long long count) {
    GRID_STRIDE(i, count) {
        const int t = seg[i >> 10];
        if (t < 0) continue;
        const float sc = (!mask || mask[t] != 0.f) ? scale[t] : 1.f;
        const float v = alpha * sc * x[i];
        y[i] = accumulate ? y[i] + v : v;
    }
}

}  // namespace

// float4 groups are legal when the row is 16-byte aligned and the 4-padded channel count fits in the row: lanes past
// the channel count may be READ (padding or a neighbouring slice, never used as data) but are never written (st4)
static inline bool mi_vec_ok(const void* p, int ld, int c) {
    return mi_al16(p) && (ld % 4 == 0) && (((c + 3) & ~3) <= ld);
}

#define LAUNCH(kernel, work, stream, ...)                                      \
    do {                                                                       \
        kernel<<<grid_for(work), TPB, 0, mi_cs(stream)>>>(__VA_ARGS__);        \
        MI_LAUNCHED();                                                         \
        MI_RETURN_LAST();                                                      \
    } while (0)

// 32-bit index arithmetic whenever the work fits (always, for the tensors of this path): the 64-bit divisions of the
// index decomposition were most of the instructions of the resampling kernels
#define LAUNCH_IDX(kernel, work, stream, ...)                                           \
    do {                                                                                \
        if ((long long)(work) < (1LL << 31))                                            \
            kernel<unsigned><<<grid_for(work), TPB, 0, mi_cs(stream)>>>(__VA_ARGS__);   \
        else                                                                            \
            kernel<long long><<<grid_for(work), TPB, 0, mi_cs(stream)>>>(__VA_ARGS__);  \
        MI_LAUNCHED();                                                                  \
        MI_RETURN_LAST();                                                               \
    } while (0)

static inline int mi_rnd_bit(int round_tf32) { return (round_tf32 && mi_tf32_rn_enabled()) ? 2 : 0; }

static bool up_strip_enabled() {
    static int on = -1;
    if (on < 0) { const char* e = getenv("MI_B200_UPSAMPLE_STRIP"); on = (e && e[0] == '0') ? 0 : 1; }
    return on == 1;
}
// strip launches of the x2 resampling kernels: `cols` x `rows` positions walked per image (output pixels forward, input
// pixels backward); false = geometry outside the grid limits, the caller launches the per-pixel form
static bool up_strip_grid(int n, int rows, int cols, int c, long long image_elems, dim3& grid, int& strip) {
    const long long per_row = (long long)cols * ((c + 3) / 4);
    static int forced = -1;                                  // MI_B200_UPSAMPLE_ROWS=1..8: fixed height (tuning aid)
    if (forced < 0) { const char* e = getenv("MI_B200_UPSAMPLE_ROWS"); forced = e ? atoi(e) : 0; }
    // as many rows per thread as leave at least half of the chip's thread slots (148 x 2048) filled
    strip = UP_STRIP_MAX;
    while (strip > 1 && per_row * rows * n / strip < 148LL * 1024) strip >>= 1;
    if (forced >= 1 && forced <= UP_STRIP_MAX) strip = forced;
    const long long gx = (per_row + UP_TPB - 1) / UP_TPB, gy = (rows + strip - 1) / strip;
    if (!up_strip_enabled() || n < 1 || n > 65535 || gy < 1 || gy > 65535 || gx < 1 || gx > 0x7fffffffLL / UP_TPB ||
        image_elems >= (1LL << 31))
        return false;
    grid = dim3((unsigned)gx, (unsigned)gy, (unsigned)n);
    return true;
}
// `vec` = VEC | RND << 1 as the per-pixel kernels take it
#define LAUNCH_UP_STRIP(kernel, vec, grid, stream, ...)                                                     \
    do {                                                                                                    \
        switch ((vec) & 3) {                                                                                \
            case 0: kernel<false, false><<<grid, UP_TPB, 0, mi_cs(stream)>>>(__VA_ARGS__); break;           \
            case 1: kernel<true, false><<<grid, UP_TPB, 0, mi_cs(stream)>>>(__VA_ARGS__); break;            \
            case 2: kernel<false, true><<<grid, UP_TPB, 0, mi_cs(stream)>>>(__VA_ARGS__); break;            \
            default: kernel<true, true><<<grid, UP_TPB, 0, mi_cs(stream)>>>(__VA_ARGS__); break;            \
        }                                                                                                   \
        MI_LAUNCHED();                                                                                      \
        MI_RETURN_LAST();                                                                                   \
    } while (0)

extern "C" {

int mi_avgpool2_fwd(const float* x, int ldx, float* y, int ldy, int n, int h, int wd, int c, int round_tf32,
                    mi_stream_t s) {
    if (!x || !y || (h & 1) || (wd & 1)) return MI_ERR_BAD_ARG;
    const int vec = (mi_vec_ok(x, ldx, c) && mi_vec_ok(y, ldy, c)) | mi_rnd_bit(round_tf32);
    LAUNCH_IDX(avgpool2_fwd_kernel, (long long)n * (h / 2) * (wd / 2) * ((c + 3) / 4), s, x, ldx, y, ldy, n, h, wd, c, vec);
}
int mi_avgpool2_bwd(const float* dy, int lddy, float* dx, int lddx, int accumulate, int n, int h, int wd, int c,
                    mi_stream_t s) {
    if (!dy || !dx || (h & 1) || (wd & 1)) return MI_ERR_BAD_ARG;
    const int vec = mi_vec_ok(dy, lddy, c) && mi_vec_ok(dx, lddx, c);
    LAUNCH_IDX(avgpool2_bwd_kernel, (long long)n * h * wd * ((c + 3) / 4), s, dy, lddy, dx, lddx, accumulate, n, h, wd, c, vec);
}
int mi_maxpool2_fwd(const float* x, int ldx, float* y, int ldy, int n, int h, int wd, int c, mi_stream_t s) {
    if (!x || !y || (h & 1) || (wd & 1)) return MI_ERR_BAD_ARG;
    LAUNCH(maxpool2_fwd_kernel, (long long)n * (h / 2) * (wd / 2) * c, s, x, ldx, y, ldy, n, h, wd, c);
}
int mi_maxpool2_bwd(const float* x, int ldx, const float* dy, int lddy, float* dx, int lddx, int accumulate, int n,
                    int h, int wd, int c, mi_stream_t s) {
    if (!x || !dy || !dx || (h & 1) || (wd & 1)) return MI_ERR_BAD_ARG;
    LAUNCH(maxpool2_bwd_kernel, (long long)n * (h / 2) * (wd / 2) * c, s, x, ldx, dy, lddy, dx, lddx, accumulate, n, h,
           wd, c);
}
int mi_upsample2_fwd(const float* x, int ldx, float* y, int ldy, int n, int h, int wd, int c, int align,
                     int round_tf32, mi_stream_t s) {
    if (!x || !y) return MI_ERR_BAD_ARG;
    const int vec = (mi_vec_ok(x, ldx, c) && mi_vec_ok(y, ldy, c)) | mi_rnd_bit(round_tf32);
    const UpWin g = {h, wd, 0, 0, 0, 0, 2 * h, 2 * wd, up2_scale(h), up2_scale(wd)};
    dim3 grid; int strip;
    if (up_strip_grid(n, g.oh, g.ow, c, std::max((long long)g.oh * g.ow * ldy, (long long)h * wd * ldx), grid, strip))
        LAUNCH_UP_STRIP(upsample2_fwd_strip_kernel, vec, grid, s, x, ldx, y, ldy, h, wd, c, align, g, strip);
    LAUNCH_IDX(upsample2_fwd_kernel, (long long)n * h * wd * 4 * ((c + 3) / 4), s, x, ldx, y, ldy, n, h, wd, c, align, vec, g);
}
int mi_upsample2_bwd(const float* dy, int lddy, float* dx, int lddx, int accumulate, int n, int h, int wd, int c,
                     int align, int round_tf32, mi_stream_t s) {
    if (!dy || !dx) return MI_ERR_BAD_ARG;
    const int vec = (mi_vec_ok(dy, lddy, c) && mi_vec_ok(dx, lddx, c)) | mi_rnd_bit(round_tf32);
    const UpWin g = {h, wd, 0, 0, 0, 0, 2 * h, 2 * wd, up2_scale(h), up2_scale(wd)};
    dim3 grid; int strip;
    if (up_strip_grid(n, h, wd, c, std::max((long long)g.oh * g.ow * lddy, (long long)h * wd * lddx), grid, strip))
        LAUNCH_UP_STRIP(upsample2_bwd_strip_kernel, vec, grid, s, dy, lddy, dx, lddx, accumulate, h, wd, c, align, g,
                        nullptr, 0, 0, 0.f, strip);
    LAUNCH_IDX(upsample2_bwd_kernel, (long long)n * h * wd * ((c + 3) / 4), s, dy, lddy, dx, lddx, accumulate, n, h, wd, c, align,
           vec, g, nullptr, 0, 0, 0.f);
}
static bool up_window_ok(int h, int wd, int full_h, int full_w, int ly0, int lx0, int oh, int ow, int hy0, int hx0) {
    return h >= 1 && wd >= 1 && oh >= 1 && ow >= 1 && ly0 >= 0 && lx0 >= 0 && hy0 >= 0 && hx0 >= 0 &&
           ly0 + h <= full_h && lx0 + wd <= full_w && hy0 + oh <= 2 * full_h && hx0 + ow <= 2 * full_w;
}
int mi_upsample2_window_fwd(const float* x, int ldx, float* y, int ldy, int n, int h, int wd, int c, int align,
                            int full_h, int full_w, int ly0, int lx0, int oh, int ow, int hy0, int hx0, int round_tf32,
                            mi_stream_t s) {
    if (!x || !y || !up_window_ok(h, wd, full_h, full_w, ly0, lx0, oh, ow, hy0, hx0)) return MI_ERR_BAD_ARG;
    const int vec = (mi_vec_ok(x, ldx, c) && mi_vec_ok(y, ldy, c)) | mi_rnd_bit(round_tf32);
    const UpWin g = {full_h, full_w, ly0, lx0, hy0, hx0, oh, ow, up2_scale(full_h), up2_scale(full_w)};
    dim3 grid; int strip;
    if (up_strip_grid(n, oh, ow, c, std::max((long long)oh * ow * ldy, (long long)h * wd * ldx), grid, strip))
        LAUNCH_UP_STRIP(upsample2_fwd_strip_kernel, vec, grid, s, x, ldx, y, ldy, h, wd, c, align, g, strip);
    LAUNCH_IDX(upsample2_fwd_kernel, (long long)n * oh * ow * ((c + 3) / 4), s, x, ldx, y, ldy, n, h, wd, c, align, vec, g);
}
int mi_upsample2_window_bwd(const float* dy, int lddy, float* dx, int lddx, int accumulate, int n, int h, int wd, int c,
                            int align, int full_h, int full_w, int ly0, int lx0, int oh, int ow, int hy0, int hx0,
                            const float* mask_y, int ldmask, int mask_act, float mask_slope, int round_tf32,
                            mi_stream_t s) {
    if (!dy || !dx || !up_window_ok(h, wd, full_h, full_w, ly0, lx0, oh, ow, hy0, hx0) || (mask_y && ldmask < c))
        return MI_ERR_BAD_ARG;
    const int vec = (mi_vec_ok(dy, lddy, c) && mi_vec_ok(dx, lddx, c) && (!mask_y || mi_vec_ok(mask_y, ldmask, c))) |
                    mi_rnd_bit(round_tf32);
    const UpWin g = {full_h, full_w, ly0, lx0, hy0, hx0, oh, ow, up2_scale(full_h), up2_scale(full_w)};
    dim3 grid; int strip;
    if (up_strip_grid(n, h, wd, c, std::max({(long long)oh * ow * lddy, (long long)h * wd * lddx, (long long)h * wd * ldmask}),
                      grid, strip))
        LAUNCH_UP_STRIP(upsample2_bwd_strip_kernel, vec, grid, s, dy, lddy, dx, lddx, accumulate, h, wd, c, align, g,
                        mask_y, ldmask, mask_act, mask_slope, strip);
    LAUNCH_IDX(upsample2_bwd_kernel, (long long)n * h * wd * ((c + 3) / 4), s, dy, lddy, dx, lddx, accumulate, n, h, wd, c, align,
           vec, g, mask_y, ldmask, mask_act, mask_slope);
}
int mi_window_copy(const float* src, int lds, int sh, int sw, int sy0, int sx0, float* dst, int ldd, int dh, int dw,
                   int dy0, int dx0, int n, int h, int wd, int c, int accumulate, mi_stream_t s) {
    if (!src || !dst || n < 1 || h < 1 || wd < 1 || c < 1 || sy0 < 0 || sx0 < 0 || dy0 < 0 || dx0 < 0 ||
        sy0 + h > sh || sx0 + wd > sw || dy0 + h > dh || dx0 + wd > dw)
        return MI_ERR_BAD_ARG;
    const int vec = mi_vec_ok(src, lds, c) && mi_vec_ok(dst, ldd, c);
    LAUNCH_IDX(window_copy_kernel, (long long)n * h * wd * ((c + 3) / 4), s, src, lds, sh, sw, sy0, sx0, dst, ldd, dh, dw, dy0,
           dx0, n, h, wd, c, accumulate, vec);
}
int mi_add(const float* a, int lda, const float* b, int ldb, float* y, int ldy, size_t pixels, int c, int round_tf32,
           mi_stream_t s) {
    if (!a || !b || !y) return MI_ERR_BAD_ARG;
    const int vec = (mi_vec_ok(a, lda, c) && mi_vec_ok(b, ldb, c) && mi_vec_ok(y, ldy, c)) | mi_rnd_bit(round_tf32);
    LAUNCH_IDX(add_kernel, (long long)pixels * ((c + 3) / 4), s, a, lda, b, ldb, y, ldy, (long long)pixels, c, vec);
}
int mi_copy(const float* src, int lds, float* dst, int ldd, int accumulate, size_t pixels, int c, mi_stream_t s) {
    if (!src || !dst) return MI_ERR_BAD_ARG;
    const int vec = mi_vec_ok(src, lds, c) && mi_vec_ok(dst, ldd, c);
    LAUNCH_IDX(copy_kernel, (long long)pixels * ((c + 3) / 4), s, src, lds, dst, ldd, accumulate, (long long)pixels, c, vec);
}
int mi_act_bwd(float* dy, int lddy, const float* y, int ldy, int act, float slope, size_t pixels, int c,
               int round_tf32, mi_stream_t s) {
    if (!dy || !y) return MI_ERR_BAD_ARG;
    if (act == MI_ACT_NONE && !round_tf32) return MI_OK;
    const int vec = mi_vec_ok(dy, lddy, c) && mi_vec_ok(y, ldy, c);
    LAUNCH_IDX(act_bwd_kernel, (long long)pixels * ((c + 3) / 4), s, dy, lddy, y, ldy, act, slope, (long long)pixels, c, vec,
           round_tf32 && mi_tf32_rn_enabled());
}
int mi_fill(float* p, float v, size_t count, mi_stream_t s) {
    if (!p) return MI_ERR_BAD_ARG;
    if (count == 0) return MI_OK;
    LAUNCH(fill_kernel, (long long)count, s, p, v, (long long)count);
}
int mi_axpby(const float* x, float a, float* y, float b, size_t count, mi_stream_t s) {
    if (!x || !y) return MI_ERR_BAD_ARG;
    if (count == 0) return MI_OK;
    LAUNCH(axpby_kernel, (long long)count, s, x, a, y, b, (long long)count);
}
int mi_addcmul(float* y, float a, const float* x1, const float* x2, size_t count, mi_stream_t s) {
    if (!y || !x1 || !x2) return MI_ERR_BAD_ARG;
    if (count == 0) return MI_OK;
    LAUNCH(addcmul_kernel, (long long)count, s, y, a, x1, x2, (long long)count);
}
int mi_frames_to_canvas(const float* f0, const float* f1, float* canvas, int ldc, int n, int h, int wd, int ch, int cw,
                        int pad_top, int pad_left, int mode, int round_tf32, mi_stream_t s) {
    if (!f0 || !f1 || !canvas || ldc < 6) return MI_ERR_BAD_ARG;
    LAUNCH(frames_to_canvas_kernel, (long long)n * ch * cw, s, f0, f1, canvas, ldc, n, h, wd, ch, cw, pad_top,
           pad_left, mode, mi_rnd_bit(round_tf32) ? 1 : 0);
}
int mi_nhwc_window_to_nchw(const float* src, int lds, float* dst, int n, int hs, int ws, int y0, int x0, int h, int wd,
                           int c, mi_stream_t s) {
    if (!src || !dst || y0 < 0 || x0 < 0 || y0 + h > hs || x0 + wd > ws) return MI_ERR_BAD_ARG;
    LAUNCH(nhwc_window_to_nchw_kernel, (long long)n * c * h * wd, s, src, lds, dst, n, hs, ws, y0, x0, h, wd, c);
}
int mi_nchw_to_nhwc_window(const float* src, float* dst, int ldd, int n, int hs, int ws, int y0, int x0, int h, int wd,
                           int c, mi_stream_t s) {
    if (!src || !dst || y0 < 0 || x0 < 0 || y0 + h > hs || x0 + wd > ws) return MI_ERR_BAD_ARG;
    LAUNCH(nchw_to_nhwc_window_kernel, (long long)n * c * h * wd, s, src, dst, ldd, n, hs, ws, y0, x0, h, wd, c);
}
int mi_loss_fwd_bwd(const float* pred, const float* target, float* grad, float* loss_out, size_t count, int kind,
                    float weight, mi_stream_t s) {
    if (!pred || !target || !loss_out || count == 0 || (kind != 0 && kind != 1)) return MI_ERR_BAD_ARG;
    LAUNCH(loss_fwd_bwd_kernel, (long long)count, s, pred, target, grad, loss_out, (long long)count, kind, weight);
}
int mi_psnr_accumulate(const float* pred, const float* target, double* sq_out, size_t count, mi_stream_t s) {
    if (!pred || !target || !sq_out) return MI_ERR_BAD_ARG;
    LAUNCH(psnr_accumulate_kernel, (long long)count, s, pred, target, sq_out, (long long)count);
}
int mi_ssim_accumulate(const float* pred, const float* target, double* sum_out, int c, int h, int w,
                       const float* window_host, int win, float val_range, mi_stream_t s) {
    if (!pred || !target || !sum_out || !window_host || win < 1 || win > SS_MAXWIN || h < win || w < win || c < 1)
        return MI_ERR_BAD_ARG;
    SsimWindow wnd;
    for (int i = 0; i < SS_MAXWIN; ++i) wnd.g[i] = i < win ? window_host[i] : 0.f;
    const float c1 = (float)((0.01 * val_range) * (0.01 * val_range));
    const float c2 = (float)((0.03 * val_range) * (0.03 * val_range));
    dim3 grid(mi_cdiv(w - win + 1, SS_TX), mi_cdiv(h - win + 1, SS_TY), c);
    ssim_accumulate_kernel<<<grid, SS_TX * SS_TY, 0, mi_cs(s)>>>(pred, target, sum_out, h, w, win, wnd, c1, c2);
    MI_LAUNCHED();
    MI_RETURN_LAST();
}
int mi_inner_update(const float* w_in, const float* g, float* w_out, float* exp_avg, float* exp_avg_sq,
                    const float* lr, int lr_per_element, int lr_stride, int num_step, const int32_t* seg,
                    const uint8_t* skip, size_t count, int rule, int step_count, mi_stream_t s) {
    if (!w_in || !g || !w_out || !lr || !seg || rule < 0 || rule > 3) return MI_ERR_BAD_ARG;
    if ((rule == 1 && (!exp_avg || !exp_avg_sq)) || (rule == 2 && !exp_avg)) return MI_ERR_BAD_ARG;
    // bias corrections in double on the host, as the reference computes them in Python floats
    const float bc1 = (float)(1.0 - pow(0.9, (double)step_count));
    const float bc2s = (float)sqrt(1.0 - pow(0.99, (double)step_count));
    LAUNCH(inner_update_kernel, (long long)count, s, w_in, g, w_out, exp_avg, exp_avg_sq, lr, lr_per_element,
           lr_stride, num_step, seg, skip, (long long)count, rule, bc1, bc2s);
}
int mi_outer_step(float* p, const float* g, float* m, float* v, size_t count, int kind, float lr, float beta1,
                  float beta2, float eps, float weight_decay, int step, mi_stream_t s) {
    if (!p || !g || kind < 0 || kind > 2 || (kind > 0 && (!m || !v))) return MI_ERR_BAD_ARG;
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const float step_size = (float)((double)lr / bc1);
    const float bc2s = (float)sqrt(1.0 - pow((double)beta2, (double)step));
    LAUNCH(outer_step_kernel, (long long)count, s, p, g, m, v, (long long)count, kind, lr, beta1, beta2, eps,
           weight_decay, step_size, bc2s);
}
int mi_segment_scale(const float* x, const float* scale, const int32_t* seg, const float* mask, float* y, float alpha,
                     int accumulate, size_t count, mi_stream_t s) {
    if (!x || !scale || !seg || !y) return MI_ERR_BAD_ARG;
    LAUNCH(segment_scale_kernel, (long long)count, s, x, scale, seg, mask, y, alpha, accumulate, (long long)count);
}
int mi_segment_dot(const float* a, const float* b, const int32_t* seg, float* out, size_t count, mi_stream_t s) {
    if (!a || !seg || !out) return MI_ERR_BAD_ARG;
    const int chunks = (int)((count + 1023) >> 10);
    segment_dot_kernel<<<chunks, TPB, 0, mi_cs(s)>>>(a, b, seg, out, (long long)count);
    MI_LAUNCHED();
    MI_RETURN_LAST();
}

}  // extern "C"
