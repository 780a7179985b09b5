// Bilinear backward warp (grid_sample) forward and flow-gradient, NHWC fp32.
//
// variant 0: SuperSloMo backWarp / RRIN warp (reference superslomo/model.py:292-302,
//            rrin/model.py:8-21): normalise 2*((x+u)/W-0.5), grid_sample(align_corners=False,
//            zeros) => sample at (x+u-0.5, y+v-0.5) (SURVEY Appx E2).
// variant 1: VoxelFlow (voxel_flow.py:471-503): coor = linspace(-1,1)[x] + s*flow,
//            grid_sample(align_corners=True, border) (SURVEY Appx E3).
// Images are data (never differentiated), so only the flow gradient exists.
// HBM-bound gather, one thread per output pixel.  The images this op sees have 3 (RGB) or 4 channels in 16-byte
// NHWC pixels, so each of the four corners is ONE 16-byte read-only load (the pad lane of a 3-channel pixel is read
// and ignored), the flow is one 8-byte load and a 4-channel result one 16-byte store; neighbouring pixels' corners
// overlap (the flow is smooth), so a warp's 128 corner loads fall into a handful of 128-byte lines that L1 serves once.
// Compulsory traffic: 16 B image + 8-16 B flow + 12-16 B result per pixel; other channel counts
// or unaligned views take the scalar path.
#include "mi_common.cuh"

namespace {

struct Sample {
    int x0, y0, x1, y1;       // corner indices (may be out of range for zeros padding)
    float tx, ty;             // fractional weights
    float gmx, gmy;           // d(ix)/d(flow_x), d(iy)/d(flow_y) (0 where clipped)
};

__device__ __forceinline__ float linspace_pm1(int i, int nsz) {
    if (nsz <= 1) return -1.f;
    const float step = 2.f / (float)(nsz - 1);
    return (i < nsz / 2) ? (-1.f + step * (float)i) : (1.f - step * (float)(nsz - 1 - i));
}

__device__ __forceinline__ Sample make_sample(int x, int y, float u, float v, int wd, int h, int variant, float sx,
                                              float sy) {
    Sample s;
    float ix, iy;
    if (variant == 0) {
        const float gx = 2.f * (((float)x + sx * u) / (float)wd - 0.5f);
        const float gy = 2.f * (((float)y + sy * v) / (float)h - 0.5f);
        ix = ((gx + 1.f) * (float)wd - 1.f) * 0.5f;
        iy = ((gy + 1.f) * (float)h - 1.f) * 0.5f;
        s.gmx = sx;  // (2/W) * (W/2)
        s.gmy = sy;
    } else {
        const float gx = linspace_pm1(x, wd) + sx * u;
        const float gy = linspace_pm1(y, h) + sy * v;
        ix = (gx + 1.f) * 0.5f * (float)(wd - 1);
        iy = (gy + 1.f) * 0.5f * (float)(h - 1);
        s.gmx = sx * 0.5f * (float)(wd - 1);
        s.gmy = sy * 0.5f * (float)(h - 1);
        // border padding: clip and kill the gradient outside (ATen clip_coordinates_set_grad)
        if (ix <= 0.f) { ix = 0.f; s.gmx = 0.f; } else if (ix >= (float)(wd - 1)) { ix = (float)(wd - 1); s.gmx = 0.f; }
        if (iy <= 0.f) { iy = 0.f; s.gmy = 0.f; } else if (iy >= (float)(h - 1)) { iy = (float)(h - 1); s.gmy = 0.f; }
    }
    const float fx = floorf(ix), fy = floorf(iy);
    s.x0 = (int)fx; s.y0 = (int)fy; s.x1 = s.x0 + 1; s.y1 = s.y0 + 1;
    s.tx = ix - fx; s.ty = iy - fy;
    return s;
}

__device__ __forceinline__ bool inb(int x, int y, int wd, int h) { return x >= 0 && x < wd && y >= 0 && y < h; }

__device__ __forceinline__ float4 corner4(const float* __restrict__ b, int x, int y, int wd, int h) {
    return inb(x, y, wd, h) ? __ldg(reinterpret_cast<const float4*>(b + ((long long)y * wd + x) * 4))
                            : make_float4(0.f, 0.f, 0.f, 0.f);
}
__device__ __forceinline__ float2 flow2(const float* __restrict__ flow, long long i, int ldfl, bool vec) {
    if (vec) return __ldg(reinterpret_cast<const float2*>(flow + i * ldfl));
    return make_float2(flow[i * ldfl], flow[i * ldfl + 1]);
}

// 16-byte pixels (ldi == 4, c <= 4): vector loads
__global__ void __launch_bounds__(256)
warp_fwd_vec_kernel(const float* __restrict__ img, const float* __restrict__ flow, int ldfl, float* __restrict__ out,
                    int ldo, int n, int h, int wd, int c, int variant, float sx, float sy, int flow_vec, int out_vec) {
    const long long total = (long long)n * h * wd;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        long long p = i;
        const int x = (int)(p % wd); p /= wd;
        const int y = (int)(p % h);
        const int nn = (int)(p / h);
        const float2 uv = flow2(flow, i, ldfl, flow_vec);
        const Sample s = make_sample(x, y, uv.x, uv.y, wd, h, variant, sx, sy);
        const float* b = img + (long long)nn * h * wd * 4;
        const float4 v00 = corner4(b, s.x0, s.y0, wd, h), v01 = corner4(b, s.x1, s.y0, wd, h);
        const float4 v10 = corner4(b, s.x0, s.y1, wd, h), v11 = corner4(b, s.x1, s.y1, wd, h);
        const float w00 = (1.f - s.tx) * (1.f - s.ty), w01 = s.tx * (1.f - s.ty);
        const float w10 = (1.f - s.tx) * s.ty, w11 = s.tx * s.ty;
        // same order of additions as the scalar path (zero-padded corners add exact zeros)
        float4 o;
        o.x = ((0.f + w00 * v00.x) + w01 * v01.x + w10 * v10.x) + w11 * v11.x;
        o.y = ((0.f + w00 * v00.y) + w01 * v01.y + w10 * v10.y) + w11 * v11.y;
        o.z = ((0.f + w00 * v00.z) + w01 * v01.z + w10 * v10.z) + w11 * v11.z;
        o.w = ((0.f + w00 * v00.w) + w01 * v01.w + w10 * v10.w) + w11 * v11.w;
        float* d = out + i * ldo;
        if (out_vec) {
            *reinterpret_cast<float4*>(d) = o;
        } else {
            d[0] = o.x;
            if (c > 1) d[1] = o.y;
            if (c > 2) d[2] = o.z;
            if (c > 3) d[3] = o.w;
        }
    }
}

__global__ void __launch_bounds__(256)
warp_bwd_vec_kernel(const float* __restrict__ img, const float* __restrict__ flow, int ldfl,
                    const float* __restrict__ go, int ldgo, float* __restrict__ gflow, int ldgf, int accumulate, int n,
                    int h, int wd, int c, int variant, float sx, float sy, int flow_vec, int go_vec, int gf_vec) {
    const long long total = (long long)n * h * wd;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        long long p = i;
        const int x = (int)(p % wd); p /= wd;
        const int y = (int)(p % h);
        const int nn = (int)(p / h);
        const float2 uv = flow2(flow, i, ldfl, flow_vec);
        const Sample s = make_sample(x, y, uv.x, uv.y, wd, h, variant, sx, sy);
        const float* b = img + (long long)nn * h * wd * 4;
        const float4 v00 = corner4(b, s.x0, s.y0, wd, h), v01 = corner4(b, s.x1, s.y0, wd, h);
        const float4 v10 = corner4(b, s.x0, s.y1, wd, h), v11 = corner4(b, s.x1, s.y1, wd, h);
        float g[4];
        if (go_vec) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(go + i * ldgo));
            g[0] = t.x; g[1] = t.y; g[2] = t.z; g[3] = t.w;
        } else {
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) g[cc] = cc < c ? go[i * ldgo + cc] : 0.f;
        }
        const float a00[4] = {v00.x, v00.y, v00.z, v00.w}, a01[4] = {v01.x, v01.y, v01.z, v01.w};
        const float a10[4] = {v10.x, v10.y, v10.z, v10.w}, a11[4] = {v11.x, v11.y, v11.z, v11.w};
        float gx = 0.f, gy = 0.f;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            if (cc < c) {
                gx += g[cc] * ((1.f - s.ty) * (a01[cc] - a00[cc]) + s.ty * (a11[cc] - a10[cc]));
                gy += g[cc] * ((1.f - s.tx) * (a10[cc] - a00[cc]) + s.tx * (a11[cc] - a01[cc]));
            }
        }
        gx *= s.gmx;
        gy *= s.gmy;
        float* d = gflow + i * ldgf;
        if (gf_vec) {
            float2 o = make_float2(gx, gy);
            if (accumulate) { const float2 q = *reinterpret_cast<const float2*>(d); o.x += q.x; o.y += q.y; }
            *reinterpret_cast<float2*>(d) = o;
        } else {
            d[0] = accumulate ? d[0] + gx : gx;
            d[1] = accumulate ? d[1] + gy : gy;
        }
    }
}

__global__ void warp_fwd_kernel(const float* __restrict__ img, int ldi, const float* __restrict__ flow, int ldfl,
                                float* __restrict__ out, int ldo, int n, int h, int wd, int c, int variant, float sx,
                                float sy) {
    const long long total = (long long)n * h * wd;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        long long p = i;
        const int x = (int)(p % wd); p /= wd;
        const int y = (int)(p % h);
        const int nn = (int)(p / h);
        const Sample s = make_sample(x, y, flow[i * ldfl], flow[i * ldfl + 1], wd, h, variant, sx, sy);
        const float* b = img + (long long)nn * h * wd * ldi;
        const bool k00 = inb(s.x0, s.y0, wd, h), k01 = inb(s.x1, s.y0, wd, h);
        const bool k10 = inb(s.x0, s.y1, wd, h), k11 = inb(s.x1, s.y1, wd, h);
        const float w00 = (1.f - s.tx) * (1.f - s.ty), w01 = s.tx * (1.f - s.ty);
        const float w10 = (1.f - s.tx) * s.ty, w11 = s.tx * s.ty;
        for (int cc = 0; cc < c; ++cc) {
            float v = 0.f;
            if (k00) v += w00 * b[((long long)s.y0 * wd + s.x0) * ldi + cc];
            if (k01) v += w01 * b[((long long)s.y0 * wd + s.x1) * ldi + cc];
            if (k10) v += w10 * b[((long long)s.y1 * wd + s.x0) * ldi + cc];
            if (k11) v += w11 * b[((long long)s.y1 * wd + s.x1) * ldi + cc];
            out[i * ldo + cc] = v;
        }
    }
}

__global__ void warp_bwd_kernel(const float* __restrict__ img, int ldi, const float* __restrict__ flow, int ldfl,
                                const float* __restrict__ go, int ldgo, float* __restrict__ gflow, int ldgf,
                                int accumulate, int n, int h, int wd, int c, int variant, float sx, float sy) {
    const long long total = (long long)n * h * wd;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        long long p = i;
        const int x = (int)(p % wd); p /= wd;
        const int y = (int)(p % h);
        const int nn = (int)(p / h);
        const Sample s = make_sample(x, y, flow[i * ldfl], flow[i * ldfl + 1], wd, h, variant, sx, sy);
        const float* b = img + (long long)nn * h * wd * ldi;
        const bool k00 = inb(s.x0, s.y0, wd, h), k01 = inb(s.x1, s.y0, wd, h);
        const bool k10 = inb(s.x0, s.y1, wd, h), k11 = inb(s.x1, s.y1, wd, h);
        float gx = 0.f, gy = 0.f;
        for (int cc = 0; cc < c; ++cc) {
            const float v00 = k00 ? b[((long long)s.y0 * wd + s.x0) * ldi + cc] : 0.f;
            const float v01 = k01 ? b[((long long)s.y0 * wd + s.x1) * ldi + cc] : 0.f;
            const float v10 = k10 ? b[((long long)s.y1 * wd + s.x0) * ldi + cc] : 0.f;
            const float v11 = k11 ? b[((long long)s.y1 * wd + s.x1) * ldi + cc] : 0.f;
            const float g = go[i * ldgo + cc];
            gx += g * ((1.f - s.ty) * (v01 - v00) + s.ty * (v11 - v10));
            gy += g * ((1.f - s.tx) * (v10 - v00) + s.tx * (v11 - v01));
        }
        gx *= s.gmx;
        gy *= s.gmy;
        float* d = gflow + i * ldgf;
        d[0] = accumulate ? d[0] + gx : gx;
        d[1] = accumulate ? d[1] + gy : gy;
    }
}

// blocks of 256 threads that are resident at once on the whole device
int resident_wave(const void* kernel) {
    int per_sm = 0, dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 256, 0) != cudaSuccess || per_sm < 1) per_sm = 4;
    return per_sm * sms;
}

}  // namespace

extern "C" {

int mi_warp_fwd(const float* img, int ldi, const float* flow, int ldfl, float* out, int ldo, int n, int h, int wd,
                int c, int variant, float sx, float sy, mi_stream_t stream) {
    if (!img || !flow || !out || ldfl < 2 || ldi < c || ldo < c || (variant != 0 && variant != 1))
        return MI_ERR_BAD_ARG;
    const long long total = (long long)n * h * wd;
    int blocks = mi_cdiv(total, 256);
    if (ldi == 4 && c <= 4 && mi_al16(img)) {
        // exactly one resident wave (ncu: with 8 blocks per SM requested and 6 resident, the 2-block second wave cost a
        // quarter of the run time); grid-stride beyond it
        const int wave = resident_wave((const void*)warp_fwd_vec_kernel);
        if (blocks > wave) blocks = wave;
        const int flow_vec = (ldfl % 2 == 0) && ((reinterpret_cast<uintptr_t>(flow) & 7) == 0);
        const int out_vec = (c == 4) && (ldo % 4 == 0) && mi_al16(out);
        warp_fwd_vec_kernel<<<blocks, 256, 0, mi_cs(stream)>>>(img, flow, ldfl, out, ldo, n, h, wd, c, variant, sx, sy,
                                                               flow_vec, out_vec);
    } else {
        if (blocks > 148 * 8) blocks = 148 * 8;
        warp_fwd_kernel<<<blocks, 256, 0, mi_cs(stream)>>>(img, ldi, flow, ldfl, out, ldo, n, h, wd, c, variant, sx, sy);
    }
    MI_LAUNCHED();
    MI_RETURN_LAST();
}

int mi_warp_bwd(const float* img, int ldi, const float* flow, int ldfl, const float* grad_out, int ldgo,
                float* grad_flow, int ldgf, float* grad_img, int ldgi, int accumulate, int n, int h, int wd, int c,
                int variant, float sx, float sy, mi_stream_t stream) {
    (void)ldgi;
    if (!img || !flow || !grad_out || !grad_flow || ldfl < 2 || ldgf < 2 || (variant != 0 && variant != 1))
        return MI_ERR_BAD_ARG;
    if (grad_img) return MI_ERR_UNSUPPORTED;  // warped images are data on this path
    const long long total = (long long)n * h * wd;
    int blocks = mi_cdiv(total, 256);
    if (ldi == 4 && c <= 4 && mi_al16(img)) {
        const int wave = resident_wave((const void*)warp_bwd_vec_kernel);
        if (blocks > wave) blocks = wave;
        const int flow_vec = (ldfl % 2 == 0) && ((reinterpret_cast<uintptr_t>(flow) & 7) == 0);
        const int go_vec = (ldgo % 4 == 0) && mi_al16(grad_out);
        const int gf_vec = (ldgf % 2 == 0) && ((reinterpret_cast<uintptr_t>(grad_flow) & 7) == 0);
        warp_bwd_vec_kernel<<<blocks, 256, 0, mi_cs(stream)>>>(img, flow, ldfl, grad_out, ldgo, grad_flow, ldgf,
                                                               accumulate, n, h, wd, c, variant, sx, sy, flow_vec,
                                                               go_vec, gf_vec);
    } else {
        if (blocks > 148 * 8) blocks = 148 * 8;
        warp_bwd_kernel<<<blocks, 256, 0, mi_cs(stream)>>>(img, ldi, flow, ldfl, grad_out, ldgo, grad_flow, ldgf,
                                                           accumulate, n, h, wd, c, variant, sx, sy);
    }
    MI_LAUNCHED();
    MI_RETURN_LAST();
}

}  // extern "C"
