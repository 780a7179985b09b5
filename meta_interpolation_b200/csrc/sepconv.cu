// Adaptive separable convolution (SepConv local filtering), forward and the
// fused gradVertical/gradHorizontal backward.
//
// Replaces the reference's cupy-JIT kernels (sepconv/sepconv_op/sepconv.py:5-30,
// 138-190).  Those run one thread per output with 2601 taps x 3 global loads;
// here the 51x51 window of a 32x8 pixel tile is staged once in shared memory
// (replicate border folded into the staging, so the doubly padded frame of
// sepconv/model.py:244-245,261-263 is never materialised), the per-pixel
// horizontal filter lives in registers and the sum is factorised as
//   out = sum_fy v[fy] * (sum_fx in[y+fy][x+fx] * h[fx])          (SURVEY Appx E1)
// The op has no GEMM structure (filters differ per pixel) and is bound by the
// FP32 FMA pipe / shared-memory bandwidth, not HBM.
#include <cstdlib>

#include "mi_common.cuh"
#include "sepconv_quad.cuh"

namespace {

// MI_B200_SEPCONV_QUAD=0 selects the first-generation kernels below (A/B measurements)
bool use_quad() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("MI_B200_SEPCONV_QUAD");
        on = (e && e[0] == '0') ? 0 : 1;
    }
    return on == 1;
}

// threads per block of the layout transposes around the quad kernels (MI_B200_SEPCONV_TPOSE_NT=64/128/256: tuning aid)
int tpose_nt() {
    static int nt = 0;
    if (!nt) {
        const char* e = getenv("MI_B200_SEPCONV_TPOSE_NT");
        const int v = e ? atoi(e) : 0;
        nt = (v == 64 || v == 128 || v == 256) ? v : quad::TPOSE_NT;
    }
    return nt;
}

// float4 form of the transposes: NHWC rows of exactly 52 floats, planar rows a multiple of 4 pixels, aligned bases,
// 32-bit offsets inside an image's planes (MI_B200_SEPCONV_TPOSE_VEC=0: scalar form)
bool tpose_vec(const float* a, const float* b, int ld, const float* pa, const float* pb, int oh, int ow) {
    static int on = -1;
    if (on < 0) { const char* e = getenv("MI_B200_SEPCONV_TPOSE_VEC"); on = (e && e[0] == '0') ? 0 : 1; }
    return on && ld == 52 && ow % 4 == 0 && mi_al16(a) && mi_al16(b) && mi_al16(pa) && mi_al16(pb) &&
           51LL * oh * ow < (1LL << 31);
}

constexpr int TX = 32;
constexpr int TY = 8;

template <int F>
struct Geo {
    static constexpr int WIN_W = TX + F - 1;
    static constexpr int WIN_H = TY + F - 1;
    static constexpr int PITCH = WIN_W + 1;  // odd pitch
};

template <int F>
__device__ __forceinline__ void stage_window(float* smem, const float* __restrict__ frame, int c, int fh, int fw,
                                             int n_idx, int y_base, int x_base) {
    constexpr int WW = Geo<F>::WIN_W, WH = Geo<F>::WIN_H, P = Geo<F>::PITCH;
    const long long plane = (long long)fh * fw;
    const float* fb = frame + (long long)n_idx * c * plane;
    for (int i = threadIdx.x; i < c * WH * WW; i += blockDim.x) {
        const int col = i % WW;
        const int r = (i / WW) % WH;
        const int cc = i / (WW * WH);
        int sy = y_base + r, sx = x_base + col;
        sy = min(max(sy, 0), fh - 1);
        sx = min(max(sx, 0), fw - 1);
        smem[(cc * WH + r) * P + col] = fb[cc * plane + (long long)sy * fw + sx];
    }
}

// Forward.  The inner sum reads one staged value per FMA, so with one pixel per thread the kernel is bound by the
// shared-memory pipe (one 32-lane LDS per cycle per SM against four FMA issue slots; ncu: LSU wavefronts 53 %, FMA 32 %).
// Each thread therefore owns TWO vertically adjacent pixels (y, y+1): their windows share 50 of 51 rows, so every
// staged value feeds two FMAs (one per pixel, each against its own register-resident horizontal filter), halving
// the shared-memory traffic per output.  A block of 32 x 8 threads covers a 32 x 16 pixel tile.
constexpr int FWD_PY = 2;

template <int F>
struct GeoF {
    static constexpr int WIN_W = TX + F - 1;
    static constexpr int WIN_H = TY * FWD_PY + F - 1;
    static constexpr int PITCH = WIN_W + 1;
};

template <int F, int C, int MINB>
__global__ void __launch_bounds__(TX* TY, MINB)
sepconv_fwd_kernel(const float* __restrict__ frame, const float* __restrict__ vert, const float* __restrict__ horiz,
                   int ldf, float* __restrict__ out, int fh, int fw, int gh, int gw, int oh, int ow, int gy0, int gx0,
                   int iy0, int ix0) {
    extern __shared__ float smem[];
    constexpr int WW = GeoF<F>::WIN_W, WH = GeoF<F>::WIN_H, P = GeoF<F>::PITCH;
    const int n_idx = blockIdx.z;
    const int ty = threadIdx.x / TX, tx = threadIdx.x % TX;
    const int oy_base = blockIdx.y * (TY * FWD_PY), ox_base = blockIdx.x * TX;
    {   // stage the (16+50) x (32+50) window of every channel; replicate border folded in
        const long long plane = (long long)fh * fw;
        const float* fb = frame + (long long)n_idx * C * plane;
        for (int i = threadIdx.x; i < C * WH * WW; i += blockDim.x) {
            const int col = i % WW;
            const int r = (i / WW) % WH;
            const int cc = i / (WW * WH);
            const int sy = min(max(oy_base + iy0 + r, 0), fh - 1);
            const int sx = min(max(ox_base + ix0 + col, 0), fw - 1);
            smem[(cc * WH + r) * P + col] = fb[cc * plane + (long long)sy * fw + sx];
        }
    }
    __syncthreads();
    const int oy = oy_base + FWD_PY * ty, ox = ox_base + tx;
    if (oy >= oh || ox >= ow) return;
    const bool two = oy + 1 < oh;                          // the second pixel of the pair exists
    const long long gpix0 = ((long long)n_idx * gh + gy0 + oy) * gw + gx0 + ox;
    const long long gpix1 = two ? gpix0 + gw : gpix0;      // (a missing partner re-reads pixel 0; result discarded)
    const float* hp0 = horiz + gpix0 * ldf;
    const float* hp1 = horiz + gpix1 * ldf;
    const float* vp0 = vert + gpix0 * ldf;
    const float* vp1 = vert + gpix1 * ldf;
    // Per-pixel filters are rows of an NHWC tensor (one pixel = 51 taps = 208 B), so a warp-wide scalar load touches
    // 32 different lines and costs the LSU pipe as much as 32 staged-window loads.  Rows are 16-byte aligned
    // (ldf % 4 == 0 checked by the host), so taps are fetched four at a time: 13 + 13 LDG.128 instead of 102 LDG.32
    // for the horizontal filters, one LDG.128 per pixel per four window rows for the vertical ones.
    float h0[F + 1], h1[F + 1];          // tap 51 is the pad lane of the 52-float row: loaded, never used
#pragma unroll
    for (int g = 0; g < (F + 3) / 4; ++g) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(hp0) + g);
        const float4 b = __ldg(reinterpret_cast<const float4*>(hp1) + g);
        h0[4 * g] = a.x; h1[4 * g] = b.x;
        if (4 * g + 1 <= F) { h0[4 * g + 1] = a.y; h1[4 * g + 1] = b.y; }
        if (4 * g + 2 <= F) { h0[4 * g + 2] = a.z; h1[4 * g + 2] = b.z; }
        if (4 * g + 3 <= F) { h0[4 * g + 3] = a.w; h1[4 * g + 3] = b.w; }
    }
    float acc0[C], acc1[C];
#pragma unroll
    for (int cc = 0; cc < C; ++cc) { acc0[cc] = 0.f; acc1[cc] = 0.f; }
    // window row r (relative to pixel 0) is tap fy = r of pixel 0 and tap fy = r - 1 of pixel 1
    float v1_prev = 0.f;                 // tap 4q-1 of pixel 1, carried from the previous group
#pragma unroll 1
    for (int q = 0; q < (F + 1 + 3) / 4; ++q) {
        const float4 va = __ldg(reinterpret_cast<const float4*>(vp0) + q);   // taps 4q..4q+3 of pixel 0
        const float4 vb = __ldg(reinterpret_cast<const float4*>(vp1) + q);   // taps 4q..4q+3 of pixel 1
        const float v0s[4] = {va.x, va.y, va.z, va.w};
        const float v1s[4] = {v1_prev, vb.x, vb.y, vb.z};
        v1_prev = vb.w;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = 4 * q + i;
            if (r > F) break;
            const float v0 = r < F ? v0s[i] : 0.f;       // (tap 51 of the row is the pad lane)
            const float v1 = r > 0 ? v1s[i] : 0.f;
            // channels innermost: 2*C independent accumulation chains and C independent loads per tap
            const float* row = smem + (FWD_PY * ty + r) * P + tx;
            float t0[C], t1[C];
#pragma unroll
            for (int cc = 0; cc < C; ++cc) { t0[cc] = 0.f; t1[cc] = 0.f; }
#pragma unroll
            for (int f = 0; f < F; ++f) {
#pragma unroll
                for (int cc = 0; cc < C; ++cc) {
                    const float in = row[cc * WH * P + f];
                    t0[cc] = fmaf(in, h0[f], t0[cc]);
                    t1[cc] = fmaf(in, h1[f], t1[cc]);
                }
            }
#pragma unroll
            for (int cc = 0; cc < C; ++cc) {
                acc0[cc] = fmaf(v0, t0[cc], acc0[cc]);
                acc1[cc] = fmaf(v1, t1[cc], acc1[cc]);
            }
        }
    }
#pragma unroll
    for (int cc = 0; cc < C; ++cc) {
        float* o = out + (((long long)n_idx * C + cc) * oh + oy) * ow + ox;
        o[0] = acc0[cc];
        if (two) o[ow] = acc1[cc];
    }
}

template <int F, int C, int MINB>
__global__ void __launch_bounds__(TX* TY, MINB)
sepconv_bwd_kernel(const float* __restrict__ frame, const float* __restrict__ vert, const float* __restrict__ horiz,
                   int ldf, const float* __restrict__ grad_out, float* __restrict__ g_vert,
                   float* __restrict__ g_horiz, int ldg, int fh, int fw, int gh, int gw, int oh, int ow, int gy0,
                   int gx0, int iy0, int ix0, int rnd) {
    extern __shared__ float smem[];
    constexpr int WH = Geo<F>::WIN_H, P = Geo<F>::PITCH;
    const int n_idx = blockIdx.z;
    const int ty = threadIdx.x / TX, tx = threadIdx.x % TX;
    const int oy_base = blockIdx.y * TY, ox_base = blockIdx.x * TX;
    stage_window<F>(smem, frame, C, fh, fw, n_idx, oy_base + iy0, ox_base + ix0);
    __syncthreads();
    const int oy = oy_base + ty, ox = ox_base + tx;
    if (oy >= oh || ox >= ow) return;
    const long long gpix = ((long long)n_idx * gh + gy0 + oy) * gw + gx0 + ox;
    const float* hp = horiz + gpix * ldf;
    const float* vp = vert + gpix * ldf;
    // filters and their gradients move four taps at a time (rows are 16-byte aligned, see the forward kernel); the
    // fourth lane of the last group is the pad lane of the 52-float row
    float hreg[F + 1], gh_acc[F + 1];
#pragma unroll
    for (int g = 0; g < (F + 3) / 4; ++g) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(hp) + g);
        hreg[4 * g] = a.x;
        if (4 * g + 1 <= F) hreg[4 * g + 1] = a.y;
        if (4 * g + 2 <= F) hreg[4 * g + 2] = a.z;
        if (4 * g + 3 <= F) hreg[4 * g + 3] = a.w;
    }
#pragma unroll
    for (int f = 0; f <= F; ++f) gh_acc[f] = 0.f;
    float go[C];
#pragma unroll
    for (int cc = 0; cc < C; ++cc) go[cc] = grad_out[(((long long)n_idx * C + cc) * oh + oy) * ow + ox];
    float* gvp = g_vert + gpix * ldg;
#pragma unroll 1
    for (int q = 0; q < (F + 3) / 4; ++q) {
        const float4 va = __ldg(reinterpret_cast<const float4*>(vp) + q);
        const float vs[4] = {va.x, va.y, va.z, va.w};
        float gvs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int fy = 4 * q + i;
            if (fy >= F) break;
            const float vv = vs[i];
            // channels innermost: C independent loads and C independent `t` chains per tap (see the forward kernel)
            const float* row = smem + (ty + fy) * P + tx;
            float gvc[C], t[C];
#pragma unroll
            for (int cc = 0; cc < C; ++cc) { gvc[cc] = go[cc] * vv; t[cc] = 0.f; }
#pragma unroll
            for (int f = 0; f < F; ++f) {
                float in[C];
#pragma unroll
                for (int cc = 0; cc < C; ++cc) in[cc] = row[cc * WH * P + f];
                float gh = gh_acc[f];
#pragma unroll
                for (int cc = 0; cc < C; ++cc) {
                    t[cc] = fmaf(in[cc], hreg[f], t[cc]);
                    gh = fmaf(gvc[cc], in[cc], gh);
                }
                gh_acc[f] = gh;
            }
            float gv = 0.f;
#pragma unroll
            for (int cc = 0; cc < C; ++cc) gv = fmaf(go[cc], t[cc], gv);
            gvs[i] = rnd ? mi_rn_tf32(gv) : gv;      // (the filter gradients are the next dgrad / wgrad operands)
        }
        if (4 * q + 4 <= F) {
            reinterpret_cast<float4*>(gvp)[q] = make_float4(gvs[0], gvs[1], gvs[2], gvs[3]);
        } else {                             // last group: the lane past tap 50 is not ours to write
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (4 * q + i < F) gvp[4 * q + i] = gvs[i];
        }
    }
    float* ghp = g_horiz + gpix * ldg;
    if (rnd) {
#pragma unroll
        for (int f = 0; f < F; ++f) gh_acc[f] = mi_rn_tf32(gh_acc[f]);
    }
#pragma unroll
    for (int g = 0; g < F / 4; ++g)
        reinterpret_cast<float4*>(ghp)[g] =
            make_float4(gh_acc[4 * g], gh_acc[4 * g + 1], gh_acc[4 * g + 2], gh_acc[4 * g + 3]);
#pragma unroll
    for (int f = (F / 4) * 4; f < F; ++f) ghp[f] = gh_acc[f];
}

// any filter size / channel count: one thread per output pixel straight from global memory
__global__ void sepconv_fwd_generic_kernel(const float* __restrict__ frame, const float* __restrict__ vert,
                                           const float* __restrict__ horiz, int ldf, float* __restrict__ out, int n,
                                           int c, int fh, int fw, int gh, int gw, int oh, int ow, int gy0, int gx0,
                                           int iy0, int ix0, int taps) {
    const long long total = (long long)n * c * oh * ow;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        long long p = i;
        const int ox = (int)(p % ow); p /= ow;
        const int oy = (int)(p % oh); p /= oh;
        const int cc = (int)(p % c);
        const int nn = (int)(p / c);
        const long long gpix = ((long long)nn * gh + gy0 + oy) * gw + gx0 + ox;
        const float* fb = frame + ((long long)nn * c + cc) * fh * fw;
        float acc = 0.f;
        for (int fy = 0; fy < taps; ++fy) {
            const int sy = min(max(oy + iy0 + fy, 0), fh - 1);
            float t = 0.f;
            for (int fx = 0; fx < taps; ++fx) {
                const int sx = min(max(ox + ix0 + fx, 0), fw - 1);
                t = fmaf(fb[(long long)sy * fw + sx], horiz[gpix * ldf + fx], t);
            }
            acc = fmaf(vert[gpix * ldf + fy], t, acc);
        }
        out[i] = acc;
    }
}

__global__ void sepconv_bwd_generic_kernel(const float* __restrict__ frame, const float* __restrict__ vert,
                                           const float* __restrict__ horiz, int ldf,
                                           const float* __restrict__ grad_out, float* __restrict__ g_vert,
                                           float* __restrict__ g_horiz, int ldg, int n, int c, int fh, int fw, int gh,
                                           int gw, int oh, int ow, int gy0, int gx0, int iy0, int ix0, int taps,
                                           int rnd) {
    // one thread per (pixel, tap index k): computes g_vert[k] and g_horiz[k]
    const long long total = (long long)n * oh * ow * taps;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        long long p = i;
        const int kidx = (int)(p % taps); p /= taps;
        const int ox = (int)(p % ow); p /= ow;
        const int oy = (int)(p % oh);
        const int nn = (int)(p / oh);
        const long long gpix = ((long long)nn * gh + gy0 + oy) * gw + gx0 + ox;
        float av = 0.f, ah = 0.f;
        for (int cc = 0; cc < c; ++cc) {
            const float* fb = frame + ((long long)nn * c + cc) * fh * fw;
            const float go = grad_out[(((long long)nn * c + cc) * oh + oy) * ow + ox];
            const int syk = min(max(oy + iy0 + kidx, 0), fh - 1);
            const int sxk = min(max(ox + ix0 + kidx, 0), fw - 1);
            float tv = 0.f, th = 0.f;
            for (int j = 0; j < taps; ++j) {
                const int sxj = min(max(ox + ix0 + j, 0), fw - 1);
                const int syj = min(max(oy + iy0 + j, 0), fh - 1);
                tv = fmaf(fb[(long long)syk * fw + sxj], horiz[gpix * ldf + j], tv);
                th = fmaf(fb[(long long)syj * fw + sxk], vert[gpix * ldf + j], th);
            }
            av = fmaf(go, tv, av);
            ah = fmaf(go, th, ah);
        }
        g_vert[gpix * ldg + kidx] = rnd ? mi_rn_tf32(av) : av;
        g_horiz[gpix * ldg + kidx] = rnd ? mi_rn_tf32(ah) : ah;
    }
}

template <int F>
size_t smem_bytes(int c) { return (size_t)c * Geo<F>::WIN_H * Geo<F>::PITCH * sizeof(float); }

bool args_ok(const void* a, const void* b, const void* c, int n, int ch, int fh, int fw, int gh, int gw, int oh,
             int ow, int gy0, int gx0, int taps, int ld) {
    return a && b && c && n > 0 && ch > 0 && fh > 0 && fw > 0 && oh > 0 && ow > 0 && taps > 0 && ld >= taps &&
           gy0 >= 0 && gx0 >= 0 && gy0 + oh <= gh && gx0 + ow <= gw;
}

// Register budget of the two kernels: MINB = 2 caps them at 128 registers (two 256-thread blocks per SM), MINB = 1
// lets ptxas keep more staged-window loads in flight at one block per SM.  MI_B200_SEPCONV_MINB="<fwd><bwd>" (e.g.
// "21") overrides the measured defaults.
int variant_minb(int which) {
    static int v[2] = {0, 0};
    if (!v[0]) {
        v[0] = 1; v[1] = 1;   // measured on B200 (tools/bench_sepconv.py): fwd 201 vs 221 us, bwd 333 vs 460 us
        const char* e = getenv("MI_B200_SEPCONV_MINB");
        if (e && (e[0] == '1' || e[0] == '2') && (e[1] == '1' || e[1] == '2')) { v[0] = e[0] - '0'; v[1] = e[1] - '0'; }
    }
    return v[which];
}

}  // namespace

extern "C" {

size_t mi_sepconv_planar_bytes(int n, int oh, int ow, int taps) {
    return (size_t)2 * n * taps * oh * ow * sizeof(float);
}

int mi_sepconv_fwd(const float* frame, const float* vert, const float* horiz, int ldf, float* out, int n, int c,
                   int fh, int fw, int gh, int gw, int oh, int ow, int gy0, int gx0, int iy0, int ix0, int taps,
                   float* planar, mi_stream_t stream) {
    if (!args_ok(frame, vert, horiz, n, c, fh, fw, gh, gw, oh, ow, gy0, gx0, taps, ldf) || !out) return MI_ERR_BAD_ARG;
    cudaStream_t st = mi_cs(stream);
    if (taps == 51 && c == 3 && planar && use_quad()) {
        static bool attr_set = false;
        const size_t sm = quad::smem_bytes<51, 3>();
        if (!attr_set) {
            cudaError_t e = cudaFuncSetAttribute(quad::sepconv_fwd_quad_kernel<51, 3>,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
            if (e != cudaSuccess) return (int)e;
            attr_set = true;
        }
        float* vpl = planar;
        float* hpl = planar + (size_t)n * 51 * oh * ow;
        const double px = (double)n * oh * ow;
        mi_prof_begin(MI_TAG_SEPCONV_FWD, 2.0 * px * (3 * 51 * 51 + 3 * 51), 4.0 * px * (2 * 51 + 3 + 3), st);
        if (tpose_vec(vert, horiz, ldf, vpl, hpl, oh, ow))
            quad::filters_to_planar_kernel<51, true><<<dim3(mi_cdiv(ow, 32), oh, 2 * n), tpose_nt(), 0, st>>>(
                vert, horiz, ldf, vpl, hpl, gh, gw, gy0, gx0, oh, ow);
        else
            quad::filters_to_planar_kernel<51, false><<<dim3(mi_cdiv(ow, 32), oh, 2 * n), tpose_nt(), 0, st>>>(
                vert, horiz, ldf, vpl, hpl, gh, gw, gy0, gx0, oh, ow);
        MI_LAUNCHED();
        const quad::Args qa = {fh, fw, oh, ow, iy0, ix0};
        dim3 grid(mi_cdiv(ow, quad::BX), mi_cdiv(oh, quad::BY), n);
        quad::sepconv_fwd_quad_kernel<51, 3><<<grid, quad::NT, sm, st>>>(frame, vpl, hpl, out, qa);
        mi_prof_end(st);
    } else if (taps == 51 && c == 3 && ldf >= 52 && (ldf & 3) == 0 && mi_al16(vert) && mi_al16(horiz)) {
        static bool attr_set = false;
        const size_t sm = (size_t)3 * GeoF<51>::WIN_H * GeoF<51>::PITCH * sizeof(float);
        if (!attr_set) {
            cudaError_t e = cudaFuncSetAttribute(sepconv_fwd_kernel<51, 3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)sm);
            if (e == cudaSuccess)
                e = cudaFuncSetAttribute(sepconv_fwd_kernel<51, 3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
            if (e != cudaSuccess) return (int)e;
            attr_set = true;
        }
        dim3 grid(mi_cdiv(ow, TX), mi_cdiv(oh, TY * FWD_PY), n);
        const double px = (double)n * oh * ow;
        mi_prof_begin(MI_TAG_SEPCONV_FWD, 2.0 * px * (3 * 51 * 51 + 3 * 51), 4.0 * px * (2 * 51 + 3 + 3), st);
        if (variant_minb(0) == 1)
            sepconv_fwd_kernel<51, 3, 1><<<grid, TX * TY, sm, st>>>(frame, vert, horiz, ldf, out, fh, fw, gh, gw, oh, ow,
                                                                    gy0, gx0, iy0, ix0);
        else
            sepconv_fwd_kernel<51, 3, 2><<<grid, TX * TY, sm, st>>>(frame, vert, horiz, ldf, out, fh, fw, gh, gw, oh, ow,
                                                                    gy0, gx0, iy0, ix0);
        mi_prof_end(st);
    } else {
        const long long total = (long long)n * c * oh * ow;
        int blocks = mi_cdiv(total, 128);
        if (blocks > 148 * 16) blocks = 148 * 16;
        sepconv_fwd_generic_kernel<<<blocks, 128, 0, st>>>(frame, vert, horiz, ldf, out, n, c, fh, fw, gh, gw, oh, ow,
                                                           gy0, gx0, iy0, ix0, taps);
    }
    MI_LAUNCHED();
    MI_RETURN_LAST();
}

int mi_sepconv_bwd(const float* frame, const float* vert, const float* horiz, int ldf, const float* grad_out,
                   float* g_vert, float* g_horiz, int ldg, int n, int c, int fh, int fw, int gh, int gw, int oh,
                   int ow, int gy0, int gx0, int iy0, int ix0, int taps, int round_tf32, float* planar,
                   int planar_valid, float* planar_grad, mi_stream_t stream) {
    const int rnd = ((round_tf32 & 1) && mi_tf32_rn_enabled()) ? 1 : 0;
    const int zero_outside = (round_tf32 & MI_SEPCONV_ZERO_OUTSIDE) ? 1 : 0;
    if (!args_ok(frame, vert, horiz, n, c, fh, fw, gh, gw, oh, ow, gy0, gx0, taps, ldf) || !grad_out || !g_vert ||
        !g_horiz || ldg < taps)
        return MI_ERR_BAD_ARG;
    cudaStream_t st = mi_cs(stream);
    const bool quad_path = taps == 51 && c == 3 && planar && planar_grad && use_quad();
    if (zero_outside && !quad_path) {       // (the quad path zero-fills inside its last launch)
        const size_t bytes = (size_t)n * gh * gw * ldg * sizeof(float);
        cudaError_t e = cudaMemsetAsync(g_vert, 0, bytes, st);
        if (e == cudaSuccess) e = cudaMemsetAsync(g_horiz, 0, bytes, st);
        if (e != cudaSuccess) return (int)e;
    }
    if (quad_path) {
        static bool attr_set = false;
        const size_t sm = quad::smem_bytes<51, 3>();
        if (!attr_set) {
            cudaError_t e = cudaFuncSetAttribute(quad::sepconv_bwd_quad_kernel<51, 3>,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
            if (e != cudaSuccess) return (int)e;
            attr_set = true;
        }
        const size_t half = (size_t)n * 51 * oh * ow;
        float* vpl = planar;
        float* hpl = planar + half;
        float* gvpl = planar_grad;
        float* ghpl = planar_grad + half;
        const double px = (double)n * oh * ow;
        mi_prof_begin(MI_TAG_SEPCONV_BWD, 2.0 * px * (2 * 3 * 51 * 51 + 2 * 3 * 51), 4.0 * px * (4 * 51 + 3 + 3), st);
        const dim3 tgrid(mi_cdiv(ow, 32), oh, 2 * n);
        if (!planar_valid) {      // the forward of this call did not leave its planar filters behind
            if (tpose_vec(vert, horiz, ldf, vpl, hpl, oh, ow))
                quad::filters_to_planar_kernel<51, true><<<tgrid, tpose_nt(), 0, st>>>(vert, horiz, ldf, vpl, hpl, gh, gw, gy0,
                                                                                     gx0, oh, ow);
            else
                quad::filters_to_planar_kernel<51, false><<<tgrid, tpose_nt(), 0, st>>>(vert, horiz, ldf, vpl, hpl, gh, gw, gy0,
                                                                                      gx0, oh, ow);
            MI_LAUNCHED();
        }
        const quad::Args qa = {fh, fw, oh, ow, iy0, ix0};
        dim3 grid(mi_cdiv(ow, quad::BX), mi_cdiv(oh, quad::BY), 2 * n);
        quad::sepconv_bwd_quad_kernel<51, 3><<<grid, quad::NT, sm, st>>>(frame, vpl, hpl, grad_out, gvpl, ghpl, qa);
        MI_LAUNCHED();
        const dim3 pgrid(mi_cdiv(ow, 32), zero_outside ? gh : oh, 2 * n);
        if (tpose_vec(g_vert, g_horiz, ldg, gvpl, ghpl, oh, ow))
            quad::planar_to_filters_kernel<51, true><<<pgrid, tpose_nt(), 0, st>>>(gvpl, ghpl, g_vert, g_horiz, ldg, gh, gw, gy0,
                                                                                 gx0, oh, ow, rnd, zero_outside);
        else
            quad::planar_to_filters_kernel<51, false><<<pgrid, tpose_nt(), 0, st>>>(gvpl, ghpl, g_vert, g_horiz, ldg, gh, gw, gy0,
                                                                                  gx0, oh, ow, rnd, zero_outside);
        mi_prof_end(st);
    } else if (taps == 51 && c == 3 && ldf >= 52 && (ldf & 3) == 0 && ldg >= 52 && (ldg & 3) == 0 && mi_al16(vert) &&
               mi_al16(horiz) && mi_al16(g_vert) && mi_al16(g_horiz)) {
        static bool attr_set = false;
        const size_t sm = smem_bytes<51>(3);
        if (!attr_set) {
            cudaError_t e = cudaFuncSetAttribute(sepconv_bwd_kernel<51, 3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)sm);
            if (e == cudaSuccess)
                e = cudaFuncSetAttribute(sepconv_bwd_kernel<51, 3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
            if (e != cudaSuccess) return (int)e;
            attr_set = true;
        }
        dim3 grid(mi_cdiv(ow, TX), mi_cdiv(oh, TY), n);
        const double px = (double)n * oh * ow;
        mi_prof_begin(MI_TAG_SEPCONV_BWD, 2.0 * px * (2 * 3 * 51 * 51 + 2 * 3 * 51), 4.0 * px * (4 * 51 + 3 + 3), st);
        if (variant_minb(1) == 1)
            sepconv_bwd_kernel<51, 3, 1><<<grid, TX * TY, sm, st>>>(frame, vert, horiz, ldf, grad_out, g_vert, g_horiz,
                                                                    ldg, fh, fw, gh, gw, oh, ow, gy0, gx0, iy0, ix0, rnd);
        else
            sepconv_bwd_kernel<51, 3, 2><<<grid, TX * TY, sm, st>>>(frame, vert, horiz, ldf, grad_out, g_vert, g_horiz,
                                                                    ldg, fh, fw, gh, gw, oh, ow, gy0, gx0, iy0, ix0, rnd);
        mi_prof_end(st);
    } else {
        const long long total = (long long)n * oh * ow * taps;
        int blocks = mi_cdiv(total, 128);
        if (blocks > 148 * 16) blocks = 148 * 16;
        sepconv_bwd_generic_kernel<<<blocks, 128, 0, st>>>(frame, vert, horiz, ldf, grad_out, g_vert, g_horiz, ldg, n,
                                                           c, fh, fw, gh, gw, oh, ow, gy0, gx0, iy0, ix0, taps, rnd);
    }
    MI_LAUNCHED();
    MI_RETURN_LAST();
}

}  // extern "C"
