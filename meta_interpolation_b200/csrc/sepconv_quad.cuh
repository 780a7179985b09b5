// Adaptive separable convolution, second generation: four vertically adjacent pixels per thread, their FMAs issued in
// pairs (FFMA2).
//
// The first kernels (sepconv.cu) keep one pixel's whole 51-tap filter in registers and reuse every staged window
// value for two pixels: one shared-memory load per two FMAs, i.e. a ceiling of half the FP32 pipe (one 32-lane LDS
// wavefront per cycle per SM against four FMA issue slots); ncu showed 23-29 % of the FMA roof.  Here a thread owns
// the pixels (y0..y0+3, x): their windows are the same 51 columns of rows shifted by one, so a staged value feeds
// FOUR FMAs.  What one thread can hold in registers is 4 pixels x 26 taps, so every quantity is produced in two
// passes over the taps it is indexed by:
//
//   forward   out[c]  = sum_fy V[fy] * T[c][fy],  T[c][fy] = sum_fx H[fx] * in[c][y+fy][x+fx]
//             pass over fx in [0,26) then [26,51): registers hold H (4 x 26), the partial T is folded into the 12
//             output accumulators at once (the sum over fx is linear), nothing is spilled between passes
//   gradH     gH[fx]  = sum_{c,fy} (gO[c] V[fy]) * in[c][y+fy][x+fx]
//             registers hold the 4 x 26 accumulators of one half of fx, the loop runs over window rows
//   gradV     gV[fy]  = sum_{c,fx} (gO[c] H[fx]) * in[c][y+fy][x+fx]
//             the transposed walk: accumulators for one half of fy, the loop runs over window COLUMNS
//
// (the reference evaluates the same sums per output with 3 x 2601 global loads, sepconv/sepconv_op/sepconv.py:5-30,
// 138-190).  Lanes of a warp are consecutive x, so every shared-memory access -- along a row or down a column -- is a
// conflict-free 128-byte wavefront.  Per 104 FMAs a thread issues 26-29 LDS: the shared-memory pipe and the FMA pipe
// are balanced, neither waits for HBM (the 51x51x3 windows of a 32x16 tile are staged once, 66 KB, three CTAs / SM).
//
// Filter layout.  The filters arrive as NHWC conv outputs: the 51 taps of a pixel are contiguous and consecutive
// pixels are 208 bytes apart, so a warp reading "tap fy of my pixel" touches 32 different cache lines.  With one LDS
// per four FMAs left, those uncoalesced loads became the bottleneck of the first version of these kernels (measured:
// no faster than the two-pixel kernels).  The filters of the output window are therefore first transposed to
// tap-planar form [image][tap][y][x] (one coalesced pass, kept for the backward), where "tap fy of 32 consecutive
// pixels" is ONE 128-byte line; the backward writes its gradients planar as well and a second transposing pass
// returns them to NHWC (rounded to the TF32 grid on request: they are the next conv's operands).
#pragma once

namespace quad {

// Two FMAs of neighbouring pixels in ONE instruction: c0 += a0 * b, c1 += a1 * b (Blackwell's packed `fma.rn.f32x2`,
// SASS FFMA2; each half is an ordinary IEEE fma, so results are bit-identical to two FFMAs).  The staged window value
// `b` is shared by the pixels of a quad; ptxas encodes the duplicated pair as a scalar-broadcast operand
// (`FFMA2 R4, R10.F32x2.HI_LO, R114.F32, R4.F32x2.HI_LO`), and the pack / unpack moves below disappear into register
// allocation.  ncu on the scalar form: issue slots 72 % busy with the FMA pipe 50 % active -- the kernels were short of
// ISSUE slots (104 FFMA + ~28 LDS + address arithmetic per row pass), which this halves for the FMA part.
__device__ __forceinline__ void ffma2(float& c0, float& c1, float a0, float a1, float b) {
    unsigned long long ra, rb, rc, rd;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b), "f"(b));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c0), "f"(c1));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(c0), "=f"(c1) : "l"(rd));
}

constexpr int QY = 4;                    // pixels per thread
constexpr int BX = 32, BY = 16;          // pixel tile of a CTA
constexpr int NT = BX * (BY / QY);       // 128 threads
constexpr int HALF_A = 26;               // taps [0, 26) and [26, 51)

template <int F>
struct Geo {
    static constexpr int WIN_W = BX + F - 1;
    static constexpr int WIN_H = BY + F - 1;
    static constexpr int PITCH = WIN_W + 1;
};
template <int F, int C>
constexpr size_t smem_bytes() { return (size_t)C * Geo<F>::WIN_H * Geo<F>::PITCH * sizeof(float); }

template <int F, int C>
__device__ __forceinline__ void stage(float* smem, const float* __restrict__ frame, int fh, int fw, int n_idx,
                                      int y_base, int x_base) {
    // one warp per window row, lanes along it (three 128-byte wavefronts per row): no per-element division, and the
    // rows of a warp are independent loads the compiler can keep in flight together
    constexpr int WW = Geo<F>::WIN_W, WH = Geo<F>::WIN_H, P = Geo<F>::PITCH;
    const long long plane = (long long)fh * fw;
    const float* fb = frame + (long long)n_idx * C * plane;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int sx[(WW + 31) / 32];
#pragma unroll
    for (int j = 0; j < (WW + 31) / 32; ++j) sx[j] = min(max(x_base + lane + 32 * j, 0), fw - 1);   // replicate border
    // (eight window rows = 24 independent 128-byte loads per warp in flight: the CTAs of a launch start together, so
    // the staging phases of the three CTAs of an SM coincide and nothing else hides their latency)
#pragma unroll 8
    for (int rr = warp; rr < C * WH; rr += NT / 32) {
        const int cc = rr / WH, r = rr - cc * WH;
        const int sy = min(max(y_base + r, 0), fh - 1);
        const float* src = fb + cc * plane + (long long)sy * fw;
        float* dst = smem + rr * P;
#pragma unroll
        for (int j = 0; j < (WW + 31) / 32; ++j)
            if (lane + 32 * j < WW) dst[lane + 32 * j] = __ldg(src + sx[j]);
    }
}

// Planar filters of the quad owned by this thread: tap t of pixel k is base[t * plane + off[k]].
struct QuadTaps {
    const float* base;
    long long plane;
    int off[QY];
    __device__ __forceinline__ float at(int k, int t) const { return __ldg(base + (long long)t * plane + off[k]); }
};

// taps [T0, T0 + NTAP) of the four pixels -> registers
template <int T0, int NTAP>
__device__ __forceinline__ void load_taps(float (&h)[QY][HALF_A], const QuadTaps& q) {
#pragma unroll
    for (int f = 0; f < NTAP; ++f)
#pragma unroll
        for (int k = 0; k < QY; ++k) h[k][f] = q.at(k, T0 + f);
}

// One pass of the forward kernel over the horizontal taps [T0, T0 + NTAP).
template <int F, int C, int T0, int NTAP>
__device__ __forceinline__ void fwd_pass(const float* win, const QuadTaps& hp, const QuadTaps& vp,
                                         float (&acc)[QY][C]) {
    constexpr int WH = Geo<F>::WIN_H, P = Geo<F>::PITCH;
    float h[QY][HALF_A];
    load_taps<T0, NTAP>(h, hp);
    // vertical taps are fetched TWO rows ahead of their use: one row of work (~500 cycles) did not cover the latency
    // of the planar-filter loads (ncu: long-scoreboard stalls on the first use of the prefetched tap)
    float vn[QY], vnn[QY];
#pragma unroll
    for (int k = 0; k < QY; ++k) {
        vn[k] = k == 0 ? vp.at(0, 0) : 0.f;             // row 0 is tap -k of pixel k
        vnn[k] = k <= 1 ? vp.at(k, 1 - k) : 0.f;        // row 1 is tap 1 - k
    }
#pragma unroll 1
    for (int r = 0; r < F + QY - 1; ++r) {             // window row relative to the first pixel of the quad
        float v[QY];
#pragma unroll
        for (int k = 0; k < QY; ++k) {
            v[k] = vn[k];
            vn[k] = vnn[k];
            const int fy = r + 2 - k;                   // row r + 2 is tap fy of pixel k
            vnn[k] = (fy >= 0 && fy < F) ? vp.at(k, fy) : 0.f;
        }
#pragma unroll
        for (int cc = 0; cc < C; ++cc) {
            const float* row = win + (cc * WH + r) * P + T0;
            float t[QY];
#pragma unroll
            for (int k = 0; k < QY; ++k) t[k] = 0.f;
#pragma unroll
            for (int f = 0; f < NTAP; ++f) {
                const float in = row[f];
                ffma2(t[0], t[1], h[0][f], h[1][f], in);
                ffma2(t[2], t[3], h[2][f], h[3][f], in);
            }
#pragma unroll
            for (int k = 0; k < QY; ++k) acc[k][cc] = fmaf(v[k], t[k], acc[k][cc]);
        }
    }
}

struct Args {
    int fh, fw, oh, ow, iy0, ix0;
};

// planar filters [image][tap][oh][ow] of the quad (y0..y0+3, x); rows past the image reuse the last valid one
template <int F>
__device__ __forceinline__ QuadTaps quad_taps(const float* planar, const Args& a, int n_idx, int y0, int x) {
    QuadTaps q;
    q.plane = (long long)a.oh * a.ow;
    q.base = planar + (long long)n_idx * F * q.plane;
#pragma unroll
    for (int k = 0; k < QY; ++k) q.off[k] = min(y0 + k, a.oh - 1) * a.ow + x;
    return q;
}

template <int F, int C>
__global__ void __launch_bounds__(NT, 3)
sepconv_fwd_quad_kernel(const float* __restrict__ frame, const float* __restrict__ vert_pl,
                        const float* __restrict__ horiz_pl, float* __restrict__ out, Args a) {
    extern __shared__ float smem[];
    constexpr int P = Geo<F>::PITCH;
    const int n_idx = blockIdx.z;
    const int tx = threadIdx.x % BX, tq = threadIdx.x / BX;
    const int oy_base = blockIdx.y * BY, ox_base = blockIdx.x * BX;
    stage<F, C>(smem, frame, a.fh, a.fw, n_idx, oy_base + a.iy0, ox_base + a.ix0);
    __syncthreads();
    const int y0 = oy_base + QY * tq, x = ox_base + tx;
    if (y0 >= a.oh || x >= a.ow) return;
    const QuadTaps hp = quad_taps<F>(horiz_pl, a, n_idx, y0, x);
    const QuadTaps vp = quad_taps<F>(vert_pl, a, n_idx, y0, x);
    float acc[QY][C];
#pragma unroll
    for (int k = 0; k < QY; ++k)
#pragma unroll
        for (int cc = 0; cc < C; ++cc) acc[k][cc] = 0.f;
    const float* win = smem + (QY * tq) * P + tx;
    fwd_pass<F, C, 0, HALF_A>(win, hp, vp, acc);
    fwd_pass<F, C, HALF_A, F - HALF_A>(win, hp, vp, acc);
#pragma unroll
    for (int k = 0; k < QY; ++k) {
        if (y0 + k >= a.oh) break;
#pragma unroll
        for (int cc = 0; cc < C; ++cc)
            out[(((long long)n_idx * C + cc) * a.oh + y0 + k) * a.ow + x] = acc[k][cc];
    }
}

// gradHorizontal for the taps [T0, T0 + NTAP): the loop runs over window rows.
template <int F, int C, int T0, int NTAP>
__device__ __forceinline__ void gh_pass(const float* win, const QuadTaps& vp, const float (&go)[QY][C],
                                        float* gout, long long plane, const int (&goff)[QY], const bool (&ok)[QY]) {
    constexpr int WH = Geo<F>::WIN_H, P = Geo<F>::PITCH;
    float acc[QY][HALF_A];
#pragma unroll
    for (int k = 0; k < QY; ++k)
#pragma unroll
        for (int f = 0; f < NTAP; ++f) acc[k][f] = 0.f;
    float vn[QY], vnn[QY];                              // two rows ahead, as in the forward pass
#pragma unroll
    for (int k = 0; k < QY; ++k) {
        vn[k] = k == 0 ? vp.at(0, 0) : 0.f;
        vnn[k] = k <= 1 ? vp.at(k, 1 - k) : 0.f;
    }
#pragma unroll 1
    for (int r = 0; r < F + QY - 1; ++r) {
        float v[QY];
#pragma unroll
        for (int k = 0; k < QY; ++k) {
            v[k] = vn[k];
            vn[k] = vnn[k];
            const int fy = r + 2 - k;
            vnn[k] = (fy >= 0 && fy < F) ? vp.at(k, fy) : 0.f;
        }
#pragma unroll
        for (int cc = 0; cc < C; ++cc) {
            const float* row = win + (cc * WH + r) * P + T0;
            float coef[QY];
#pragma unroll
            for (int k = 0; k < QY; ++k) coef[k] = go[k][cc] * v[k];
#pragma unroll
            for (int f = 0; f < NTAP; ++f) {
                const float in = row[f];
                ffma2(acc[0][f], acc[1][f], coef[0], coef[1], in);
                ffma2(acc[2][f], acc[3][f], coef[2], coef[3], in);
            }
        }
    }
#pragma unroll
    for (int f = 0; f < NTAP; ++f)
#pragma unroll
        for (int k = 0; k < QY; ++k)
            if (ok[k]) gout[(long long)(T0 + f) * plane + goff[k]] = acc[k][f];
}

// gradVertical for the taps fy in [T0, T0 + NTAP): the loop runs over window columns (= horizontal taps).
template <int F, int C, int T0, int NTAP>
__device__ __forceinline__ void gv_pass(const float* win, const QuadTaps& hp, const float (&go)[QY][C],
                                        float* gout, long long plane, const int (&goff)[QY], const bool (&ok)[QY]) {
    constexpr int WH = Geo<F>::WIN_H, P = Geo<F>::PITCH;
    float acc[QY][HALF_A];
#pragma unroll
    for (int k = 0; k < QY; ++k)
#pragma unroll
        for (int j = 0; j < NTAP; ++j) acc[k][j] = 0.f;
    float hn[QY], hnn[QY];                              // two columns ahead
#pragma unroll
    for (int k = 0; k < QY; ++k) {
        hn[k] = hp.at(k, 0);
        hnn[k] = hp.at(k, 1);
    }
#pragma unroll 1
    for (int fx = 0; fx < F; ++fx) {
        float hk[QY];
#pragma unroll
        for (int k = 0; k < QY; ++k) {
            hk[k] = hn[k];
            hn[k] = hnn[k];
            hnn[k] = hp.at(k, min(fx + 2, F - 1));
        }
#pragma unroll
        for (int cc = 0; cc < C; ++cc) {
            // window rows T0 .. T0 + NTAP + 2 (relative to the quad's first pixel) of column x + fx
            const float* col = win + (cc * WH + T0) * P + fx;
            float coef[QY];
#pragma unroll
            for (int k = 0; k < QY; ++k) coef[k] = go[k][cc] * hk[k];
#pragma unroll
            for (int jj = 0; jj < NTAP + QY - 1; ++jj) {
                const float in = col[jj * P];
                // row T0 + jj is tap fy = T0 + jj - k of pixel k: pixels (0, 1) and (2, 3) pair up, one tap apart
#pragma unroll
                for (int kp = 0; kp < QY; kp += 2) {
                    const int j0 = jj - kp, j1 = jj - kp - 1;
                    const bool v0 = j0 >= 0 && j0 < NTAP, v1 = j1 >= 0 && j1 < NTAP;
                    if (v0 && v1) ffma2(acc[kp][v0 ? j0 : 0], acc[kp + 1][v1 ? j1 : 0], coef[kp], coef[kp + 1], in);
                    else if (v0) acc[kp][v0 ? j0 : 0] = fmaf(coef[kp], in, acc[kp][v0 ? j0 : 0]);
                    else if (v1) acc[kp + 1][v1 ? j1 : 0] = fmaf(coef[kp + 1], in, acc[kp + 1][v1 ? j1 : 0]);
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < NTAP; ++j)
#pragma unroll
        for (int k = 0; k < QY; ++k)
            if (ok[k]) gout[(long long)(T0 + j) * plane + goff[k]] = acc[k][j];
}

// blockIdx.z = 2 * image + which: which 0 -> gradHorizontal, 1 -> gradVertical (two independent halves of the work,
// each staging the same window, so one launch fills the GPU with twice the CTAs).  Filters in and gradients out are
// tap-planar [image][tap][oh][ow].
template <int F, int C>
__global__ void __launch_bounds__(NT, 3)
sepconv_bwd_quad_kernel(const float* __restrict__ frame, const float* __restrict__ vert_pl,
                        const float* __restrict__ horiz_pl, const float* __restrict__ grad_out,
                        float* __restrict__ gv_pl, float* __restrict__ gh_pl, Args a) {
    extern __shared__ float smem[];
    constexpr int P = Geo<F>::PITCH;
    const int n_idx = blockIdx.z >> 1, which = blockIdx.z & 1;
    const int tx = threadIdx.x % BX, tq = threadIdx.x / BX;
    const int oy_base = blockIdx.y * BY, ox_base = blockIdx.x * BX;
    stage<F, C>(smem, frame, a.fh, a.fw, n_idx, oy_base + a.iy0, ox_base + a.ix0);
    __syncthreads();
    const int y0 = oy_base + QY * tq, x = ox_base + tx;
    if (y0 >= a.oh || x >= a.ow) return;
    const QuadTaps fp = quad_taps<F>(which ? horiz_pl : vert_pl, a, n_idx, y0, x);   // gradV needs H, gradH needs V
    const long long plane = (long long)a.oh * a.ow;
    float* gout = (which ? gv_pl : gh_pl) + (long long)n_idx * F * plane;
    bool ok[QY];
    int goff[QY];
    float go[QY][C];
#pragma unroll
    for (int k = 0; k < QY; ++k) {
        ok[k] = y0 + k < a.oh;
        goff[k] = fp.off[k];
#pragma unroll
        for (int cc = 0; cc < C; ++cc)
            go[k][cc] = ok[k] ? __ldg(grad_out + ((long long)n_idx * C + cc) * plane + goff[k]) : 0.f;
    }
    const float* win = smem + (QY * tq) * P + tx;
    if (which == 0) {
        gh_pass<F, C, 0, HALF_A>(win, fp, go, gout, plane, goff, ok);
        gh_pass<F, C, HALF_A, F - HALF_A>(win, fp, go, gout, plane, goff, ok);
    } else {
        gv_pass<F, C, 0, HALF_A>(win, fp, go, gout, plane, goff, ok);
        gv_pass<F, C, HALF_A, F - HALF_A>(win, fp, go, gout, plane, goff, ok);
    }
}

// NHWC filters [n][gh][gw][ld] (window at (gy0, gx0)) -> tap-planar [n][F][oh][ow].  One CTA = 32 consecutive pixels
// of one row: the 32 x F block is read as one contiguous run and written as F 128-byte lines.  A block moves 6.5 KB
// in, then 6.5 KB out, with a barrier in between, so the bytes in flight per SM are set by the number of resident
// blocks: launched with quad::TPOSE_NT = 128 threads (16 blocks per SM) rather than 256 (8).
// VEC (rows of F + 1 floats, ow a multiple of 4, 16-byte aligned bases: the launcher checks): both sides move float4s.
// ncu on the scalar form: issue slots 78 % busy, 60 instructions per element (a division, 64-bit address arithmetic
// and a bounds test per 4 bytes) at 35 % of the HBM rate -- instruction-bound, not memory-bound.
constexpr int TPOSE_NT = 128;
template <int F, bool VEC>
__global__ void __launch_bounds__(256)
filters_to_planar_kernel(const float* __restrict__ src0, const float* __restrict__ src1, int ld, float* __restrict__ dst0,
                         float* __restrict__ dst1, int gh, int gw, int gy0, int gx0, int oh, int ow) {
    __shared__ float tile[32][F + 2];
    const float* src = blockIdx.z & 1 ? src1 : src0;
    float* dst = blockIdx.z & 1 ? dst1 : dst0;
    const int n_idx = blockIdx.z >> 1, y = blockIdx.y, x0 = blockIdx.x * 32;
    const int npx = min(32, ow - x0);
    const float* row = src + (((long long)n_idx * gh + gy0 + y) * gw + gx0 + x0) * ld;
    const int nt = blockDim.x;
    const long long plane = (long long)oh * ow;
    float* out = dst + (long long)n_idx * F * plane + (long long)y * ow + x0;
    if constexpr (VEC) {
        constexpr int G = (F + 1) / 4;                       // float4 groups of a pixel's row
        const float4* row4 = reinterpret_cast<const float4*>(row);
        for (int i = threadIdx.x; i < npx * G; i += nt) {
            const int px = i / G, t = 4 * (i - px * G);
            const float4 v = __ldg(row4 + i);
            tile[px][t] = v.x; tile[px][t + 1] = v.y; tile[px][t + 2] = v.z; tile[px][t + 3] = v.w;   // (t + 3 <= F: the pad lane)
        }
        __syncthreads();
        const int plane_i = (int)plane;                      // (F * plane < 2^31: the launcher checks)
        for (int i = threadIdx.x; i < F * 8; i += nt) {
            const int t = i >> 3, px = 4 * (i & 7);
            float* o = out + t * plane_i + px;
            if (px + 3 < npx) {
                *reinterpret_cast<float4*>(o) = make_float4(tile[px][t], tile[px + 1][t], tile[px + 2][t], tile[px + 3][t]);
            } else {
                for (int q = 0; px + q < npx; ++q) o[q] = tile[px + q][t];
            }
        }
    } else {
        for (int i = threadIdx.x; i < npx * ld; i += nt) {
            const int px = i / ld, t = i - px * ld;
            if (t < F) tile[px][t] = __ldg(row + i);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < F * 32; i += nt) {
            const int t = i >> 5, px = i & 31;
            if (px < npx) out[(long long)t * plane + px] = tile[px][t];
        }
    }
}

// tap-planar gradients [n][F][oh][ow] -> NHWC [n][gh][gw][ld] window (rounded to the TF32 grid on request); with
// `zero_outside` (grid.y = gh) the rest of the grid is zero-filled in the same launch -- the caller's two 48 MB fills
// of a backward pass touched every byte of the window a second time
template <int F, bool VEC>
__global__ void __launch_bounds__(256)
planar_to_filters_kernel(const float* __restrict__ src0, const float* __restrict__ src1, float* __restrict__ dst0,
                         float* __restrict__ dst1, int ld, int gh, int gw, int gy0, int gx0, int oh, int ow, int rnd,
                         int zero_outside) {
    __shared__ float tile[32][F + 2];
    const float* src = blockIdx.z & 1 ? src1 : src0;
    float* dst = blockIdx.z & 1 ? dst1 : dst0;
    const int n_idx = blockIdx.z >> 1, x0 = blockIdx.x * 32;
    int y = blockIdx.y;
    if (zero_outside) {
        // blockIdx.y walks the rows of the whole grid: rows outside the window are zeroed by the blocks of the row
        // together, the margins left / right of the window by the first / last block of a window row
        const int r = blockIdx.y;
        float* grow = dst + ((long long)n_idx * gh + r) * gw * ld;
        if (r < gy0 || r >= gy0 + oh) {
            const int total = gw * ld, per = (total + gridDim.x - 1) / gridDim.x;
            const int lo = blockIdx.x * per, hi = min(total, lo + per);
            for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) grow[i] = 0.f;
            return;
        }
        if (blockIdx.x == 0)
            for (int i = threadIdx.x; i < gx0 * ld; i += blockDim.x) grow[i] = 0.f;
        if (blockIdx.x == gridDim.x - 1)
            for (int i = (gx0 + ow) * ld + threadIdx.x; i < gw * ld; i += blockDim.x) grow[i] = 0.f;
        y = r - gy0;
    }
    const int npx = min(32, ow - x0);
    const long long plane = (long long)oh * ow;
    const float* in = src + (long long)n_idx * F * plane + (long long)y * ow + x0;
    float* row = dst + (((long long)n_idx * gh + gy0 + y) * gw + gx0 + x0) * ld;
    const int nt = blockDim.x;
    if constexpr (VEC) {
        const int plane_i = (int)plane;
        for (int i = threadIdx.x; i < F * 8; i += nt) {
            const int t = i >> 3, px = 4 * (i & 7);
            const float* p = in + t * plane_i + px;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (px + 3 < npx) {
                const float4 q = __ldg(reinterpret_cast<const float4*>(p));
                v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
            } else {
                for (int q = 0; px + q < npx; ++q) v[q] = __ldg(p + q);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) tile[px + q][t] = rnd ? mi_rn_tf32(v[q]) : v[q];
        }
        __syncthreads();
        constexpr int G = (F + 1) / 4;
        for (int i = threadIdx.x; i < npx * G; i += nt) {
            const int px = i / G, g = i - px * G, t = 4 * g;
            float* o = row + px * (F + 1) + t;
            if (t + 3 < F) {
                *reinterpret_cast<float4*>(o) = make_float4(tile[px][t], tile[px][t + 1], tile[px][t + 2], tile[px][t + 3]);
            } else if (zero_outside && t + 3 == F) {          // the caller's fresh buffer: the pad lane is zeroed too,
                *reinterpret_cast<float4*>(o) = make_float4(tile[px][t], tile[px][t + 1], tile[px][t + 2], 0.f);   // whole sectors
            } else {                                          // pad lanes of somebody's NHWC rows are never written
                for (int q = 0; t + q < F; ++q) o[q] = tile[px][t + q];
            }
        }
    } else {
        for (int i = threadIdx.x; i < F * 32; i += nt) {
            const int t = i >> 5, px = i & 31;
            if (px < npx) {
                const float v = __ldg(in + (long long)t * plane + px);
                tile[px][t] = rnd ? mi_rn_tf32(v) : v;
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < npx * ld; i += nt) {
            const int px = i / ld, t = i - px * ld;
            if (t < F) row[i] = tile[px][t];   // pad lanes of the NHWC rows are never written
        }
    }
}

}  // namespace quad
