// Input/output heads of the flow- and attention-based backbones (VoxelFlow / SuperSloMo / RRIN / CAIN):
// standalone activation, clamp, the visibility-weighted blend of two warped frames, the ring kernels that turn
// the zero-padding conv engine into a reflection-padding one, CAIN's space-to-depth / depth-to-space image
// transforms with the mean shift folded in, and the channel-attention pooling / rescaling.
// All NHWC fp32, HBM-bound, one thread per element or per pixel; reductions are deterministic (no atomics).
#include <cooperative_groups.h>

#include "mi_common.cuh"

namespace cg = cooperative_groups;

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

__device__ __forceinline__ int reflect_index(int i, int n) {
    if (n == 1) return 0;
    const int period = 2 * (n - 1);
    i = i < 0 ? -i : i;
    i %= period;
    return i < n ? i : period - i;
}

// ----------------------------------------------------------------------------- activation / clamp
__global__ void act_fwd_kernel(const float* __restrict__ x, int ldx, float* __restrict__ y, int ldy, int act,
                               float slope, long long pixels, int c) {
    const long long total = pixels * c;
    GRID_STRIDE(i, total) {
        const int ch = (int)(i % c);
        const long long p = i / c;
        y[p * ldy + ch] = mi_act_apply(x[p * ldx + ch], act, slope);
    }
}

__global__ void clamp_fwd_kernel(const float* __restrict__ x, int ldx, float* __restrict__ y, int ldy, float lo,
                                 float hi, long long pixels, int c) {
    const long long total = pixels * c;
    GRID_STRIDE(i, total) {
        const int ch = (int)(i % c);
        const long long p = i / c;
        y[p * ldy + ch] = fminf(fmaxf(x[p * ldx + ch], lo), hi);
    }
}

// torch.clamp backward: the gradient passes where lo <= x <= hi (bounds included)
__global__ void clamp_bwd_kernel(const float* __restrict__ dy, int lddy, const float* __restrict__ x, int ldx,
                                 float* __restrict__ dx, int lddx, int accumulate, float lo, float hi,
                                 long long pixels, int c) {
    const long long total = pixels * c;
    GRID_STRIDE(i, total) {
        const int ch = (int)(i % c);
        const long long p = i / c;
        const float xv = x[p * ldx + ch];
        const float g = (xv >= lo && xv <= hi) ? dy[p * lddy + ch] : 0.f;
        float* d = dx + p * lddx + ch;
        *d = accumulate ? *d + g : g;
    }
}

// ----------------------------------------------------------------------------- blend of two warped frames
// mode 0: out = (w0*m0*a + w1*m1*b) / (w0*m0 + w1*m1 + eps)          (rrin/model.py:102-103)
// mode 1: same with m1 = 1 - m0                                      (superslomo/model.py:621-630)
// mode 2: out = m0*a + (1 - m0)*b                                    (voxel_flow.py:505-507)
// a, b: c channels; m0, m1: one channel each.  One thread per pixel.
__global__ void blend_fwd_kernel(const float* __restrict__ a, int lda, const float* __restrict__ b, int ldb,
                                 const float* __restrict__ m0, int ldm0, const float* __restrict__ m1, int ldm1,
                                 float* __restrict__ out, int ldo, float w0, float w1, float eps, int mode,
                                 long long pixels, int c) {
    GRID_STRIDE(p, pixels) {
        const float v0 = m0[p * ldm0];
        const float v1 = mode == 0 ? m1[p * ldm1] : 1.f - v0;
        if (mode == 2) {
            for (int ch = 0; ch < c; ++ch) out[p * ldo + ch] = v0 * a[p * lda + ch] + v1 * b[p * ldb + ch];
        } else {
            const float ca = w0 * v0, cb = w1 * v1;
            const float den = ca + cb + eps;
            for (int ch = 0; ch < c; ++ch) out[p * ldo + ch] = (ca * a[p * lda + ch] + cb * b[p * ldb + ch]) / den;
        }
    }
}

__global__ void blend_bwd_kernel(const float* __restrict__ a, int lda, const float* __restrict__ b, int ldb,
                                 const float* __restrict__ m0, int ldm0, const float* __restrict__ m1, int ldm1,
                                 const float* __restrict__ go, int ldgo, float* __restrict__ ga, int ldga,
                                 float* __restrict__ gb, int ldgb, float* __restrict__ gm0, int ldgm0,
                                 float* __restrict__ gm1, int ldgm1, int accumulate, float w0, float w1, float eps,
                                 int mode, long long pixels, int c) {
    GRID_STRIDE(p, pixels) {
        const float v0 = m0[p * ldm0];
        const float v1 = mode == 0 ? m1[p * ldm1] : 1.f - v0;
        float s0 = 0.f, s1 = 0.f;
        if (mode == 2) {
            for (int ch = 0; ch < c; ++ch) {
                const float g = go[p * ldgo + ch];
                const float av = a[p * lda + ch], bv = b[p * ldb + ch];
                if (ga) { float* d = ga + p * ldga + ch; *d = accumulate ? *d + g * v0 : g * v0; }
                if (gb) { float* d = gb + p * ldgb + ch; *d = accumulate ? *d + g * v1 : g * v1; }
                s0 += g * (av - bv);
            }
            if (gm0) { float* d = gm0 + p * ldgm0; *d = accumulate ? *d + s0 : s0; }
        } else {
            const float ca = w0 * v0, cb = w1 * v1;
            const float inv = 1.f / (ca + cb + eps);
            for (int ch = 0; ch < c; ++ch) {
                const float g = go[p * ldgo + ch] * inv;
                const float av = a[p * lda + ch], bv = b[p * ldb + ch];
                const float o = (ca * av + cb * bv) * inv;
                if (ga) { float* d = ga + p * ldga + ch; *d = accumulate ? *d + g * ca : g * ca; }
                if (gb) { float* d = gb + p * ldgb + ch; *d = accumulate ? *d + g * cb : g * cb; }
                s0 += g * (av - o);
                s1 += g * (bv - o);
            }
            s0 *= w0;
            s1 *= w1;
            if (mode == 1) {
                if (gm0) { float* d = gm0 + p * ldgm0; *d = accumulate ? *d + (s0 - s1) : (s0 - s1); }
            } else {
                if (gm0) { float* d = gm0 + p * ldgm0; *d = accumulate ? *d + s0 : s0; }
                if (gm1) { float* d = gm1 + p * ldgm1; *d = accumulate ? *d + s1 : s1; }
            }
        }
    }
}

// ----------------------------------------------------------------------------- ring kernels (reflection padding)
// An activation of logical size (h-2) x (w-2) lives in the interior of an h x w buffer.  ring_fix fills the
// one-pixel ring with zeros (mode 0: the conv engine's own zero padding, made explicit) or with the reflection of
// the interior (mode 1: nn.ReflectionPad2d(1), model_utils.py:825-826), so that the zero-padding conv engine run
// over the whole buffer yields, in the interior, the reflection-padded convolution.  ring_fold is its transpose:
// the gradient that reached a ring pixel is added to the interior pixel it mirrored (mode 1) and the ring is
// cleared (both modes).
__device__ __forceinline__ void ring_pixel(long long r, int h, int w, int* y, int* x) {
    // enumerate the 2*w + 2*(h-2) ring pixels of one image
    if (r < w) { *y = 0; *x = (int)r; return; }
    r -= w;
    if (r < w) { *y = h - 1; *x = (int)r; return; }
    r -= w;
    *y = 1 + (int)(r >> 1);
    *x = (r & 1) ? w - 1 : 0;
}
__device__ __forceinline__ int ring_src(int i, int n) {   // mirrored interior index of a (possibly ring) index
    return i == 0 ? 2 : (i == n - 1 ? n - 3 : i);
}

__global__ void ring_fix_kernel(float* __restrict__ x, int ld, int n, int h, int w, int c, int mode) {
    const long long per = 2LL * w + 2LL * (h - 2);
    const long long total = (long long)n * per * c;
    GRID_STRIDE(i, total) {
        const int ch = (int)(i % c);
        long long r = i / c;
        const int img = (int)(r / per);
        r -= (long long)img * per;
        int y, xx;
        ring_pixel(r, h, w, &y, &xx);
        float v = 0.f;
        if (mode == 1) v = x[(((long long)img * h + ring_src(y, h)) * w + ring_src(xx, w)) * ld + ch];
        x[(((long long)img * h + y) * w + xx) * ld + ch] = v;
    }
}

// Reflection padding is separable (rows, then columns), so its transpose is too: fold the two ring ROWS into rows
// 2 and h-3 over every column (ring columns included), then fold the two ring COLUMNS of rows 1..h-2 into columns 2
// and w-3.  Every destination has exactly one writer per pass -- deterministic, no atomics; the ring is cleared by
// the thread that consumed it.
__global__ void ring_fold_rows_kernel(float* __restrict__ g, int ld, int n, int h, int w, int c) {
    const long long total = (long long)n * w * c;
    GRID_STRIDE(i, total) {
        const int ch = (int)(i % c);
        long long r = i / c;
        const int x = (int)(r % w);
        const int img = (int)(r / w);
        float* base = g + (long long)img * h * w * ld + (long long)x * ld + ch;
        const long long row = (long long)w * ld;
        base[2 * row] += base[0];
        base[(long long)(h - 3) * row] += base[(long long)(h - 1) * row];
        base[0] = 0.f;
        base[(long long)(h - 1) * row] = 0.f;
    }
}
__global__ void ring_fold_cols_kernel(float* __restrict__ g, int ld, int n, int h, int w, int c) {
    const long long total = (long long)n * (h - 2) * c;
    GRID_STRIDE(i, total) {
        const int ch = (int)(i % c);
        long long r = i / c;
        const int y = 1 + (int)(r % (h - 2));
        const int img = (int)(r / (h - 2));
        float* base = g + ((long long)img * h + y) * w * ld + ch;
        base[2LL * ld] += base[0];
        base[(long long)(w - 3) * ld] += base[(long long)(w - 1) * ld];
        base[0] = 0.f;
        base[(long long)(w - 1) * ld] = 0.f;
    }
}

// ----------------------------------------------------------------------------- CAIN image transforms
// per-(image, channel) mean of an NCHW tensor: mean over W of each row, then mean over H of those
// (model_utils.py:11-15 takes mean(2) then mean(3); for equal-sized rows the two orders agree to rounding).
__global__ void channel_mean_nchw_kernel(const float* __restrict__ f, float* __restrict__ out, int hw) {
    __shared__ float red[TPB];
    const float* src = f + (long long)blockIdx.x * hw;
    float s = 0.f;
    for (int i = threadIdx.x; i < hw; i += TPB) s += src[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int k = TPB / 2; k > 0; k >>= 1) {
        if ((int)threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = red[0] / (float)hw;
}

// out[n, 1+y, 1+x, f*3*r*r + c*r*r + by*r + bx] = reflect(frame_f)[n, c, y*r+by - pad_top, x*r+bx - pad_left] - mean_f[n,c]
// (cain/model.py:70-77 sub_mean + InOutPaddings, model_utils.py:202-217 with scale 1/r); the output buffer has a
// one-pixel ring, written as zeros (= the zero padding of headConv).
__global__ void space_to_depth_kernel(const float* __restrict__ f0, const float* __restrict__ f1,
                                      const float* __restrict__ mean0, const float* __restrict__ mean1,
                                      float* __restrict__ out, int ldo, int n, int h, int wd, int pad_top,
                                      int pad_left, int oh, int ow, int r) {
    const int cpf = 3 * r * r;                      // channels per frame
    const long long total = (long long)n * (oh + 2) * (ow + 2) * 2 * cpf;
    GRID_STRIDE(i, total) {
        long long p = i;
        const int ch = (int)(p % (2 * cpf)); p /= 2 * cpf;
        const int x = (int)(p % (ow + 2)); p /= ow + 2;
        const int y = (int)(p % (oh + 2));
        const int img = (int)(p / (oh + 2));
        float v = 0.f;
        if (y >= 1 && y <= oh && x >= 1 && x <= ow) {
            const int fsel = ch / cpf, cc = ch - fsel * cpf;
            const int c = cc / (r * r), rem = cc - c * r * r;
            const int by = rem / r, bx = rem - by * r;
            const int sy = reflect_index((y - 1) * r + by - pad_top, h);
            const int sx = reflect_index((x - 1) * r + bx - pad_left, wd);
            const float* f = fsel ? f1 : f0;
            const float* m = fsel ? mean1 : mean0;
            v = f[(((long long)img * 3 + c) * h + sy) * wd + sx] - m[img * 3 + c];
        }
        out[(((long long)img * (oh + 2) + y) * (ow + 2) + x) * ldo + ch] = v;
    }
}

// out[n, c, Y, X] = in[n, 1 + (Y+pad_top)/r, 1 + (X+pad_left)/r, c*r*r + ((Y+pad_top)%r)*r + (X+pad_left)%r] + (m0+m1)/2
__global__ void depth_to_space_kernel(const float* __restrict__ in, int ldi, const float* __restrict__ mean0,
                                      const float* __restrict__ mean1, float* __restrict__ out, int n, int h, int wd,
                                      int pad_top, int pad_left, int ih, int iw, int r) {
    const long long total = (long long)n * 3 * h * wd;
    GRID_STRIDE(i, total) {
        long long p = i;
        const int X = (int)(p % wd); p /= wd;
        const int Y = (int)(p % h); p /= h;
        const int c = (int)(p % 3);
        const int img = (int)(p / 3);
        const int py = Y + pad_top, px = X + pad_left;
        const int ch = c * r * r + (py % r) * r + (px % r);
        const float v = in[(((long long)img * (ih + 2) + 1 + py / r) * (iw + 2) + 1 + px / r) * ldi + ch];
        out[i] = v + 0.5f * (mean0[img * 3 + c] + mean1[img * 3 + c]);
    }
}

// transpose of the above: every element of the padded NHWC gradient buffer is written (zero outside the crop / ring)
__global__ void depth_to_space_bwd_kernel(const float* __restrict__ gout, float* __restrict__ gin, int ldi, int n,
                                          int h, int wd, int pad_top, int pad_left, int ih, int iw, int r) {
    const int cpf = 3 * r * r;
    const long long total = (long long)n * (ih + 2) * (iw + 2) * cpf;
    GRID_STRIDE(i, total) {
        long long p = i;
        const int ch = (int)(p % cpf); p /= cpf;
        const int x = (int)(p % (iw + 2)); p /= iw + 2;
        const int y = (int)(p % (ih + 2));
        const int img = (int)(p / (ih + 2));
        float v = 0.f;
        if (y >= 1 && y <= ih && x >= 1 && x <= iw) {
            const int c = ch / (r * r), rem = ch - c * r * r;
            const int by = rem / r, bx = rem - by * r;
            const int Y = (y - 1) * r + by - pad_top, X = (x - 1) * r + bx - pad_left;
            if (Y >= 0 && Y < h && X >= 0 && X < wd) v = gout[(((long long)img * 3 + c) * h + Y) * wd + X];
        }
        gin[(((long long)img * (ih + 2) + y) * (iw + 2) + x) * ldi + ch] = v;
    }
}

// ----------------------------------------------------------------------------- channel attention
// mean over the interior (ring excluded) of each channel.  One thread-block CLUSTER of eight CTAs per (image, 32-channel
// group): every CTA (32 channel lanes x 32 pixel lanes, coalesced 128-byte rows, four loads in flight per thread)
// reduces one eighth of the pixels, then CTA 0 collects the eight partial rows through distributed shared memory in
// rank order -- deterministic, no workspace, no second launch.  (The single-CTA form put 12 blocks on 148 SMs and was
// a quarter of a CAIN task.)  `mul` (optional) turns it into sum(x*mul), the gradient of the per-channel scale
// (model_utils.py:931-955: x * y).
constexpr int IR_CLUSTER = 8;

__global__ void __cluster_dims__(1, 1, IR_CLUSTER) __launch_bounds__(1024)
interior_reduce_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ mul, int ldm,
                       float* __restrict__ out, int h, int w, int c, int ring, float scale) {
    __shared__ float red[32][33];
    __shared__ float part[32];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int img = blockIdx.x;
    const int ch = blockIdx.y * 32 + threadIdx.x;
    const int ih = h - 2 * ring, iw = w - 2 * ring;
    const long long npix = (long long)ih * iw;
    const long long per = (npix + IR_CLUSTER - 1) / IR_CLUSTER;
    const long long q_end = min(npix, (long long)(rank + 1) * per);
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    if (ch < c) {
        const float* bx = x + (long long)img * h * w * ldx + ch;
        const float* bm = mul ? mul + (long long)img * h * w * ldm + ch : nullptr;
        for (long long q = (long long)rank * per + threadIdx.y; q < q_end; q += 128) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const long long qq = q + 32 * u;
                if (qq < q_end) {
                    const int y0 = (int)(qq / iw), x0 = (int)(qq - (long long)y0 * iw);
                    const long long o0 = (long long)(y0 + ring) * w + x0 + ring;
                    float a0 = bx[o0 * ldx];
                    if (bm) a0 *= bm[o0 * ldm];
                    s[u] += a0;
                }
            }
        }
    }
    red[threadIdx.y][threadIdx.x] = (s[0] + s[1]) + (s[2] + s[3]);
    __syncthreads();
    if (threadIdx.y == 0) {
        float t = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) t += red[j][threadIdx.x];
        part[threadIdx.x] = t;
    }
    cluster.sync();                                   // every CTA's partial row is in its shared memory
    if (rank == 0 && threadIdx.y == 0 && ch < c) {
        float t = 0.f;
        for (int r = 0; r < IR_CLUSTER; ++r) t += cluster.map_shared_rank(part, r)[threadIdx.x];
        out[(long long)img * c + ch] = t * scale;
    }
    cluster.sync();                                   // remote shared memory stays alive until CTA 0 has read it
}

// The same reduction with 16-byte loads for 16-byte-aligned rows: 8 lanes x float4 cover the 32 channels of a group,
// 128 pixel lanes, four independent loads per thread = 64 KB in flight per CTA (the scalar form above kept 16 KB in
// flight on 96 of the 148 SMs and ran at 1.5 TB/s: 9 % of a CAIN task).
__global__ void __cluster_dims__(1, 1, IR_CLUSTER) __launch_bounds__(1024)
interior_reduce_vec_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ mul, int ldm,
                           float* __restrict__ out, int h, int w, int c, int ring, float scale) {
    __shared__ float red[128][33];
    __shared__ float part[32];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int img = blockIdx.x;
    const int c4 = blockIdx.y * 32 + threadIdx.x * 4;          // first of this thread's four channels
    const int ih = h - 2 * ring, iw = w - 2 * ring;
    const long long npix = (long long)ih * iw;
    const long long per = (npix + IR_CLUSTER - 1) / IR_CLUSTER;
    const long long q_end = min(npix, (long long)(rank + 1) * per);
    float4 s[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) s[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c4 < c) {                                               // (c is a multiple of 4 on this path)
        const float* bx = x + (long long)img * h * w * ldx + c4;
        const float* bm = mul ? mul + (long long)img * h * w * ldm + c4 : nullptr;
        for (long long q = (long long)rank * per + threadIdx.y; q < q_end; q += 512) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const long long qq = q + 128 * u;
                if (qq < q_end) {
                    const int y0 = (int)(qq / iw), x0 = (int)(qq - (long long)y0 * iw);
                    const long long o0 = (long long)(y0 + ring) * w + x0 + ring;
                    float4 a = __ldg(reinterpret_cast<const float4*>(bx + o0 * ldx));
                    if (bm) {
                        const float4 m = __ldg(reinterpret_cast<const float4*>(bm + o0 * ldm));
                        a.x *= m.x; a.y *= m.y; a.z *= m.z; a.w *= m.w;
                    }
                    s[u].x += a.x; s[u].y += a.y; s[u].z += a.z; s[u].w += a.w;
                }
            }
        }
    }
    float* r = &red[threadIdx.y][threadIdx.x * 4];
    r[0] = (s[0].x + s[1].x) + (s[2].x + s[3].x); r[1] = (s[0].y + s[1].y) + (s[2].y + s[3].y);
    r[2] = (s[0].z + s[1].z) + (s[2].z + s[3].z); r[3] = (s[0].w + s[1].w) + (s[2].w + s[3].w);
    __syncthreads();
    const int tid = threadIdx.y * 8 + threadIdx.x;
    if (tid < 32) {
        float t = 0.f;
#pragma unroll 8
        for (int j = 0; j < 128; ++j) t += red[j][tid];
        part[tid] = t;
    }
    cluster.sync();                                   // every CTA's partial row is in its shared memory
    if (rank == 0 && tid < 32 && blockIdx.y * 32 + tid < c) {
        float t = 0.f;
        for (int rr = 0; rr < IR_CLUSTER; ++rr) t += cluster.map_shared_rank(part, rr)[tid];
        out[(long long)img * c + blockIdx.y * 32 + tid] = t * scale;
    }
    cluster.sync();                                   // remote shared memory stays alive until CTA 0 has read it
}

// out = o * s[n,c] + res   (RCAB: x * y then out += res, model_utils.py:955,985)
__global__ void scale_add_kernel(const float* __restrict__ o, int ldo, const float* __restrict__ s,
                                 const float* __restrict__ res, int ldr, float* __restrict__ out, int ldout,
                                 long long pix_per_img, long long pixels, int c) {
    const long long total = pixels * c;
    GRID_STRIDE(i, total) {
        const int ch = (int)(i % c);
        const long long p = i / c;
        const int img = (int)(p / pix_per_img);
        const float r = res ? res[p * ldr + ch] : 0.f;
        out[p * ldout + ch] = o[p * ldo + ch] * s[(long long)img * c + ch] + r;
    }
}

// dx (+)= g * s[n,c]  (gradient of the rescale w.r.t. its feature input)
__global__ void scale_bwd_kernel(const float* __restrict__ g, int ldg, const float* __restrict__ s,
                                 float* __restrict__ dx, int lddx, int accumulate, long long pix_per_img,
                                 long long pixels, int c) {
    const long long total = pixels * c;
    GRID_STRIDE(i, total) {
        const int ch = (int)(i % c);
        const long long p = i / c;
        const int img = (int)(p / pix_per_img);
        const float v = g[p * ldg + ch] * s[(long long)img * c + ch];
        float* d = dx + p * lddx + ch;
        *d = accumulate ? *d + v : v;
    }
}

// dx[interior] += dy[n,c] * scale  (gradient of the interior mean)
__global__ void interior_bcast_add_kernel(const float* __restrict__ dy, float* __restrict__ dx, int lddx, int n, int h,
                                          int w, int c, int ring, float scale) {
    const int ih = h - 2 * ring, iw = w - 2 * ring;
    const long long total = (long long)n * ih * iw * c;
    GRID_STRIDE(i, total) {
        long long p = i;
        const int ch = (int)(p % c); p /= c;
        const int x = (int)(p % iw); p /= iw;
        const int y = (int)(p % ih);
        const int img = (int)(p / ih);
        dx[(((long long)img * h + y + ring) * w + x + ring) * lddx + ch] += dy[(long long)img * c + ch] * scale;
    }
}

}  // namespace

extern "C" {

int mi_act_fwd(const float* x, int ldx, float* y, int ldy, int act, float slope, size_t pixels, int c,
               mi_stream_t stream) {
    if (!x || !y) return MI_ERR_BAD_ARG;
    act_fwd_kernel<<<grid_for((long long)pixels * c), TPB, 0, mi_cs(stream)>>>(x, ldx, y, ldy, act, slope,
                                                                              (long long)pixels, c);
    MI_LAUNCHED();
    MI_RETURN_LAST();
}

int mi_clamp_fwd(const float* x, int ldx, float* y, int ldy, float lo, float hi, size_t pixels, int c,
                 mi_stream_t stream) {
    if (!x || !y) return MI_ERR_BAD_ARG;
    clamp_fwd_kernel<<<grid_for((long long)pixels * c), TPB, 0, mi_cs(stream)>>>(x, ldx, y, ldy, lo, hi,
                                                                                (long long)pixels, c);
    MI_LAUNCHED();
    MI_RETURN_LAST();
}

int mi_clamp_bwd(const float* dy, int lddy, const float* x, int ldx, float* dx, int lddx, int accumulate, float lo,
                 float hi, size_t pixels, int c, mi_stream_t stream) {
    if (!dy || !x || !dx) return MI_ERR_BAD_ARG;
    clamp_bwd_kernel<<<grid_for((long long)pixels * c), TPB, 0, mi_cs(stream)>>>(dy, lddy, x, ldx, dx, lddx, accumulate,
                                                                                lo, hi, (long long)pixels, c);
    MI_LAUNCHED();
    MI_RETURN_LAST();
}

int mi_blend_fwd(const float* a, int lda, const float* b, int ldb, const float* m0, int ldm0, const float* m1,
                 int ldm1, float* out, int ldo, float w0, float w1, float eps, int mode, size_t pixels, int c,
                 mi_stream_t stream) {
    if (!a || !b || !m0 || !out || mode < 0 || mode > 2 || (mode == 0 && !m1)) return MI_ERR_BAD_ARG;
    blend_fwd_kernel<<<grid_for((long long)pixels), TPB, 0, mi_cs(stream)>>>(a, lda, b, ldb, m0, ldm0, m1, ldm1, out, ldo,
                                                                            w0, w1, eps, mode, (long long)pixels, c);
    MI_LAUNCHED();
    MI_RETURN_LAST();
}

int mi_blend_bwd(const float* a, int lda, const float* b, int ldb, const float* m0, int ldm0, const float* m1,
                 int ldm1, const float* go, int ldgo, float* ga, int ldga, float* gb, int ldgb, float* gm0, int ldgm0,
                 float* gm1, int ldgm1, int accumulate, float w0, float w1, float eps, int mode, size_t pixels, int c,
                 mi_stream_t stream) {
    if (!a || !b || !m0 || !go || mode < 0 || mode > 2 || (mode == 0 && !m1)) return MI_ERR_BAD_ARG;
    blend_bwd_kernel<<<grid_for((long long)pixels), TPB, 0, mi_cs(stream)>>>(
        a, lda, b, ldb, m0, ldm0, m1, ldm1, go, ldgo, ga, ldga, gb, ldgb, gm0, ldgm0, gm1, ldgm1, accumulate, w0, w1, eps,
        mode, (long long)pixels, c);
    MI_LAUNCHED();
    MI_RETURN_LAST();
}

int mi_ring_fix(float* x, int ld, int n, int h, int wd, int c, int mode, mi_stream_t stream) {
    if (!x || h < 5 || wd < 5 || (mode != 0 && mode != 1)) return MI_ERR_BAD_ARG;
    const long long total = (long long)n * (2LL * wd + 2LL * (h - 2)) * c;
    ring_fix_kernel<<<grid_for(total), TPB, 0, mi_cs(stream)>>>(x, ld, n, h, wd, c, mode);
    MI_LAUNCHED();
    MI_RETURN_LAST();
}

int mi_ring_fold(float* g, int ld, int n, int h, int wd, int c, int mode, mi_stream_t stream) {
    if (!g || h < 5 || wd < 5 || (mode != 0 && mode != 1)) return MI_ERR_BAD_ARG;
    if (mode == 0) {
        const long long total = (long long)n * (2LL * wd + 2LL * (h - 2)) * c;
        ring_fix_kernel<<<grid_for(total), TPB, 0, mi_cs(stream)>>>(g, ld, n, h, wd, c, 0);
        MI_LAUNCHED();
        MI_RETURN_LAST();
    }
    ring_fold_rows_kernel<<<grid_for((long long)n * wd * c), TPB, 0, mi_cs(stream)>>>(g, ld, n, h, wd, c);
    MI_LAUNCHED();
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) return (int)e;
    ring_fold_cols_kernel<<<grid_for((long long)n * (h - 2) * c), TPB, 0, mi_cs(stream)>>>(g, ld, n, h, wd, c);
    MI_LAUNCHED();
    MI_RETURN_LAST();
}

int mi_channel_mean_nchw(const float* f, float* out, int planes, int hw, mi_stream_t stream) {
    if (!f || !out || planes < 1 || hw < 1) return MI_ERR_BAD_ARG;
    channel_mean_nchw_kernel<<<planes, TPB, 0, mi_cs(stream)>>>(f, out, hw);
    MI_LAUNCHED();
    MI_RETURN_LAST();
}

int mi_space_to_depth(const float* f0, const float* f1, const float* mean0, const float* mean1, float* out, int ldo,
                      int n, int h, int wd, int pad_top, int pad_left, int oh, int ow, int r, mi_stream_t stream) {
    if (!f0 || !f1 || !mean0 || !mean1 || !out || r < 1 || ldo < 6 * r * r) return MI_ERR_BAD_ARG;
    const long long total = (long long)n * (oh + 2) * (ow + 2) * 6 * r * r;
    space_to_depth_kernel<<<grid_for(total), TPB, 0, mi_cs(stream)>>>(f0, f1, mean0, mean1, out, ldo, n, h, wd, pad_top,
                                                                     pad_left, oh, ow, r);
    MI_LAUNCHED();
    MI_RETURN_LAST();
}

int mi_depth_to_space(const float* in, int ldi, const float* mean0, const float* mean1, float* out, int n, int h,
                      int wd, int pad_top, int pad_left, int ih, int iw, int r, mi_stream_t stream) {
    if (!in || !mean0 || !mean1 || !out || r < 1 || ldi < 3 * r * r) return MI_ERR_BAD_ARG;
    if ((h + pad_top + r - 1) / r > ih || (wd + pad_left + r - 1) / r > iw) return MI_ERR_BAD_ARG;
    depth_to_space_kernel<<<grid_for((long long)n * 3 * h * wd), TPB, 0, mi_cs(stream)>>>(in, ldi, mean0, mean1, out, n, h,
                                                                                         wd, pad_top, pad_left, ih, iw, r);
    MI_LAUNCHED();
    MI_RETURN_LAST();
}

int mi_depth_to_space_bwd(const float* gout, float* gin, int ldi, int n, int h, int wd, int pad_top, int pad_left,
                          int ih, int iw, int r, mi_stream_t stream) {
    if (!gout || !gin || r < 1 || ldi < 3 * r * r) return MI_ERR_BAD_ARG;
    const long long total = (long long)n * (ih + 2) * (iw + 2) * 3 * r * r;
    depth_to_space_bwd_kernel<<<grid_for(total), TPB, 0, mi_cs(stream)>>>(gout, gin, ldi, n, h, wd, pad_top, pad_left, ih,
                                                                         iw, r);
    MI_LAUNCHED();
    MI_RETURN_LAST();
}

int mi_interior_reduce(const float* x, int ldx, const float* mul, int ldm, float* out, int n, int h, int wd, int c,
                       int ring, float scale, mi_stream_t stream) {
    if (!x || !out || ring < 0 || h <= 2 * ring || wd <= 2 * ring) return MI_ERR_BAD_ARG;
    const bool vec = (c % 4 == 0) && (ldx % 4 == 0) && mi_al16(x) && (!mul || ((ldm % 4 == 0) && mi_al16(mul)));
    if (vec)
        interior_reduce_vec_kernel<<<dim3(n, mi_cdiv(c, 32), IR_CLUSTER), dim3(8, 128), 0, mi_cs(stream)>>>(
            x, ldx, mul, ldm, out, h, wd, c, ring, scale);
    else
        interior_reduce_kernel<<<dim3(n, mi_cdiv(c, 32), IR_CLUSTER), dim3(32, 32), 0, mi_cs(stream)>>>(x, ldx, mul, ldm, out,
                                                                                                       h, wd, c, ring, scale);
    MI_LAUNCHED();
    MI_RETURN_LAST();
}

int mi_scale_add(const float* o, int ldo, const float* s, const float* res, int ldr, float* out, int ldout, int n,
                 size_t pixels_per_image, int c, mi_stream_t stream) {
    if (!o || !s || !out) return MI_ERR_BAD_ARG;
    const long long pixels = (long long)n * (long long)pixels_per_image;
    scale_add_kernel<<<grid_for(pixels * c), TPB, 0, mi_cs(stream)>>>(o, ldo, s, res, ldr, out, ldout,
                                                                     (long long)pixels_per_image, pixels, c);
    MI_LAUNCHED();
    MI_RETURN_LAST();
}

int mi_scale_bwd(const float* g, int ldg, const float* s, float* dx, int lddx, int accumulate, int n,
                 size_t pixels_per_image, int c, mi_stream_t stream) {
    if (!g || !s || !dx) return MI_ERR_BAD_ARG;
    const long long pixels = (long long)n * (long long)pixels_per_image;
    scale_bwd_kernel<<<grid_for(pixels * c), TPB, 0, mi_cs(stream)>>>(g, ldg, s, dx, lddx, accumulate,
                                                                     (long long)pixels_per_image, pixels, c);
    MI_LAUNCHED();
    MI_RETURN_LAST();
}

int mi_interior_bcast_add(const float* dy, float* dx, int lddx, int n, int h, int wd, int c, int ring, float scale,
                          mi_stream_t stream) {
    if (!dy || !dx || ring < 0 || h <= 2 * ring || wd <= 2 * ring) return MI_ERR_BAD_ARG;
    const long long total = (long long)n * (h - 2 * ring) * (wd - 2 * ring) * c;
    interior_bcast_add_kernel<<<grid_for(total), TPB, 0, mi_cs(stream)>>>(dy, dx, lddx, n, h, wd, c, ring, scale);
    MI_LAUNCHED();
    MI_RETURN_LAST();
}

}  // extern "C"
