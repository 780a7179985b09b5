// Glue kernels of the flow-based backbones (VoxelFlow / SuperSloMo / RRIN): frozen batch-norm (+activation)
// forward/backward, broadcasting binary ops with their gradients, affine scaling.  All NHWC fp32, HBM-bound,
// one thread per (pixel, 4-channel group) like elementwise.cu.
#include "mi_common.cuh"

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

// ----------------------------------------------------------------------------- frozen batch norm (+ activation)
// y = act( (x - mean[c]) * rsqrt(var[c] + eps) * gamma[c] + beta[c] )   (nn.BatchNorm2d in eval mode,
// reference voxel_flow.py:241-263, 352-355)
__global__ void bn_eval_fwd_kernel(const float* __restrict__ x, int ldx, float* __restrict__ y, int ldy,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   const float* __restrict__ mean, const float* __restrict__ var, float eps, int act,
                                   float slope, long long pixels, int c) {
    const long long total = pixels * c;
    GRID_STRIDE(i, total) {
        const int ch = (int)(i % c);
        const long long p = i / c;
        const float inv = rsqrtf(var[ch] + eps);
        const float v = (x[p * ldx + ch] - mean[ch]) * inv * gamma[ch] + beta[ch];
        y[p * ldy + ch] = mi_act_apply(v, act, slope);
    }
}

// dz = dy * act'(y); dx = dz * gamma * inv_std;  per-slice partials of dbeta = sum dz, dgamma = sum dz * xhat.
// grid = (slices, ceil(c/32)), block = 32 channels x 8 pixel lanes.
__global__ void bn_eval_bwd_kernel(const float* __restrict__ dy, int lddy, const float* __restrict__ y, int ldy,
                                   const float* __restrict__ x, int ldx, float* __restrict__ dx, int lddx,
                                   const float* __restrict__ gamma, const float* __restrict__ mean,
                                   const float* __restrict__ var, float eps, int act, float slope,
                                   float* __restrict__ ws, long long pixels, long long chunk, int c, int accumulate_dx) {
    const int split = blockIdx.x;
    const long long m0 = (long long)split * chunk;
    long long m1 = m0 + chunk;
    if (m1 > pixels) m1 = pixels;
    const int ch = blockIdx.y * 32 + threadIdx.x;
    __shared__ float red_b[8][33], red_g[8][33];
    float sb = 0.f, sg = 0.f;
    if (ch < c) {
        const float inv = rsqrtf(var[ch] + eps), mu = mean[ch], g = gamma[ch];
        for (long long m = m0 + threadIdx.y; m < m1; m += 8) {
            const float dz = dy[m * lddy + ch] * mi_act_grad(y[m * ldy + ch], act, slope);
            const float xhat = (x[m * ldx + ch] - mu) * inv;
            sb += dz;
            sg += dz * xhat;
            if (dx) {
                float* d = dx + m * lddx + ch;
                const float v = dz * g * inv;
                *d = accumulate_dx ? *d + v : v;
            }
        }
    }
    red_b[threadIdx.y][threadIdx.x] = sb;
    red_g[threadIdx.y][threadIdx.x] = sg;
    __syncthreads();
    if (threadIdx.y == 0 && ch < c) {
        float b = 0.f, g = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) { b += red_b[j][threadIdx.x]; g += red_g[j][threadIdx.x]; }
        ws[((long long)split * 2 + 0) * c + ch] = g;
        ws[((long long)split * 2 + 1) * c + ch] = b;
    }
}

__global__ void bn_eval_finish_kernel(const float* __restrict__ ws, int splits, int c, int mode, float scale,
                                      float* __restrict__ dgamma, float* __restrict__ dbeta) {
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= c) return;
    float g = 0.f, b = 0.f;
    for (int s = 0; s < splits; ++s) {
        g += ws[((long long)s * 2 + 0) * c + ch];
        b += ws[((long long)s * 2 + 1) * c + ch];
    }
    if (mode == MI_WG_ACCUM) {
        if (dgamma) dgamma[ch] += scale * g;
        if (dbeta) dbeta[ch] += scale * b;
    } else {
        if (dgamma) dgamma[ch] = g;
        if (dbeta) dbeta[ch] = b;
    }
}

// ----------------------------------------------------------------------------- broadcasting binary ops
// y = a (op) b, where b has either the same channel count as a or a single channel (broadcast over channels).
__device__ __forceinline__ float bin_apply(int op, float a, float b) {
    switch (op) {
        case 0: return a + b;
        case 1: return a - b;
        case 2: return a * b;
        default: return a / b;
    }
}

__global__ void binary_fwd_kernel(int op, const float* __restrict__ a, int lda, const float* __restrict__ b, int ldb,
                                  int cb, float* __restrict__ y, int ldy, long long pixels, int c) {
    const long long total = pixels * c;
    GRID_STRIDE(i, total) {
        const int ch = (int)(i % c);
        const long long p = i / c;
        const float bv = b[p * ldb + (cb == 1 ? 0 : ch)];
        y[p * ldy + ch] = bin_apply(op, a[p * lda + ch], bv);
    }
}

// one thread per pixel: loops channels so the broadcast operand's gradient is a register sum (no atomics)
__global__ void binary_bwd_kernel(int op, const float* __restrict__ a, int lda, const float* __restrict__ b, int ldb,
                                  int cb, const float* __restrict__ go, int ldgo, float* __restrict__ ga, int ldga,
                                  int acc_a, float* __restrict__ gb, int ldgb, int acc_b, long long pixels, int c) {
    GRID_STRIDE(p, pixels) {
        float gsum = 0.f;
        for (int ch = 0; ch < c; ++ch) {
            const float g = go[p * ldgo + ch];
            const float av = a[p * lda + ch];
            const float bv = b[p * ldb + (cb == 1 ? 0 : ch)];
            float da, db;
            switch (op) {
                case 0: da = g; db = g; break;
                case 1: da = g; db = -g; break;
                case 2: da = g * bv; db = g * av; break;
                default: da = g / bv; db = -g * av / (bv * bv); break;
            }
            if (ga) {
                float* d = ga + p * ldga + ch;
                *d = acc_a ? *d + da : da;
            }
            if (gb) {
                if (cb == 1) {
                    gsum += db;
                } else {
                    float* d = gb + p * ldgb + ch;
                    *d = acc_b ? *d + db : db;
                }
            }
        }
        if (gb && cb == 1) {
            float* d = gb + p * ldgb;
            *d = acc_b ? *d + gsum : gsum;
        }
    }
}

// y = alpha * x + beta  (and its gradient dx (+)= alpha * dy)
__global__ void affine_kernel(const float* __restrict__ x, int ldx, float* __restrict__ y, int ldy, float alpha,
                              float beta, int accumulate, long long pixels, int c) {
    const long long total = pixels * c;
    GRID_STRIDE(i, total) {
        const int ch = (int)(i % c);
        const long long p = i / c;
        const float v = alpha * x[p * ldx + ch] + beta;
        float* d = y + p * ldy + ch;
        *d = accumulate ? *d + v : v;
    }
}

}  // namespace

extern "C" {

int mi_bn_eval_fwd(const float* x, int ldx, float* y, int ldy, const float* gamma, const float* beta,
                   const float* mean, const float* var, float eps, int act, float slope, size_t pixels, int c,
                   mi_stream_t stream) {
    if (!x || !y || !gamma || !beta || !mean || !var) return MI_ERR_BAD_ARG;
    bn_eval_fwd_kernel<<<grid_for((long long)pixels * c), TPB, 0, mi_cs(stream)>>>(
        x, ldx, y, ldy, gamma, beta, mean, var, eps, act, slope, (long long)pixels, c);
    MI_LAUNCHED();
    MI_RETURN_LAST();
}

size_t mi_bn_eval_bwd_workspace(size_t pixels, int c) {
    long long s = (long long)pixels / 1024;
    if (s < 1) s = 1;
    if (s > 128) s = 128;
    return (size_t)s * 2 * c * sizeof(float);
}

int mi_bn_eval_bwd(const float* dy, int lddy, const float* y, int ldy, const float* x, int ldx, float* dx, int lddx,
                   int accumulate_dx, const float* gamma, const float* mean, const float* var, float eps, int act,
                   float slope, float* dgamma, float* dbeta, int mode, float scale, void* workspace,
                   size_t workspace_bytes, size_t pixels, int c, mi_stream_t stream) {
    if (!dy || !y || !x || !gamma || !mean || !var || !workspace) return MI_ERR_BAD_ARG;
    if (workspace_bytes < mi_bn_eval_bwd_workspace(pixels, c)) return MI_ERR_WORKSPACE;
    long long splits = (long long)pixels / 1024;
    if (splits < 1) splits = 1;
    if (splits > 128) splits = 128;
    const long long chunk = ((long long)pixels + splits - 1) / splits;
    float* ws = reinterpret_cast<float*>(workspace);
    bn_eval_bwd_kernel<<<dim3((unsigned)splits, mi_cdiv(c, 32)), dim3(32, 8), 0, mi_cs(stream)>>>(
        dy, lddy, y, ldy, x, ldx, dx, lddx, gamma, mean, var, eps, act, slope, ws, (long long)pixels, chunk, c,
        accumulate_dx);
    MI_LAUNCHED();
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) return (int)e;
    if (dgamma || dbeta) {
        bn_eval_finish_kernel<<<mi_cdiv(c, 128), 128, 0, mi_cs(stream)>>>(ws, (int)splits, c, mode, scale, dgamma, dbeta);
        MI_LAUNCHED();
    }
    MI_RETURN_LAST();
}

int mi_binary_fwd(int op, const float* a, int lda, const float* b, int ldb, int cb, float* y, int ldy, size_t pixels,
                  int c, mi_stream_t stream) {
    if (!a || !b || !y || op < 0 || op > 3 || (cb != 1 && cb != c)) return MI_ERR_BAD_ARG;
    binary_fwd_kernel<<<grid_for((long long)pixels * c), TPB, 0, mi_cs(stream)>>>(op, a, lda, b, ldb, cb, y, ldy,
                                                                                 (long long)pixels, c);
    MI_LAUNCHED();
    MI_RETURN_LAST();
}

int mi_binary_bwd(int op, const float* a, int lda, const float* b, int ldb, int cb, const float* go, int ldgo,
                  float* ga, int ldga, int acc_a, float* gb, int ldgb, int acc_b, size_t pixels, int c,
                  mi_stream_t stream) {
    if (!a || !b || !go || op < 0 || op > 3 || (cb != 1 && cb != c)) return MI_ERR_BAD_ARG;
    binary_bwd_kernel<<<grid_for((long long)pixels), TPB, 0, mi_cs(stream)>>>(op, a, lda, b, ldb, cb, go, ldgo, ga, ldga,
                                                                            acc_a, gb, ldgb, acc_b, (long long)pixels, c);
    MI_LAUNCHED();
    MI_RETURN_LAST();
}

int mi_affine(const float* x, int ldx, float* y, int ldy, float alpha, float beta, int accumulate, size_t pixels,
              int c, mi_stream_t stream) {
    if (!x || !y) return MI_ERR_BAD_ARG;
    affine_kernel<<<grid_for((long long)pixels * c), TPB, 0, mi_cs(stream)>>>(x, ldx, y, ldy, alpha, beta, accumulate,
                                                                             (long long)pixels, c);
    MI_LAUNCHED();
    MI_RETURN_LAST();
}

}  // extern "C"
