// Shared helpers for libmi_b200 (sm_100a).  Host-side launch bookkeeping and
// small device utilities used by every translation unit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include "../../include/mi_b200.h"

extern std::atomic<unsigned long long> g_mi_launches;

#define MI_LAUNCHED() (g_mi_launches.fetch_add(1, std::memory_order_relaxed))
#define MI_RETURN_LAST()                         \
    do {                                         \
        cudaError_t e__ = cudaPeekAtLastError(); \
        return (int)e__;                         \
    } while (0)

// per-launch profiling hooks (profile.cu); no-ops unless mi_prof_enable(1)
enum { MI_TAG_FPROP_TC = 0, MI_TAG_WGRAD_TC = 1, MI_TAG_FPROP_SIMT = 2, MI_TAG_WGRAD_SIMT = 3, MI_TAG_SEPCONV_FWD = 4,
       MI_TAG_SEPCONV_BWD = 5, MI_TAG_WGRAD_FINISH = 6, MI_TAG_FPROP_HALO = 7, MI_TAG_FPROP_STREAM = 8,
       MI_TAG_WGRAD_KX = 9, MI_TAG_FPROP_KXS = 10 };
void mi_prof_begin(int tag, double flops, double bytes, cudaStream_t s);
void mi_prof_end(cudaStream_t s);
static inline double mi_conv_flops(int n, int h, int w, int cin, int cout, int k) {
    return 2.0 * n * h * w * (double)cin * cout * k * k;
}
static inline double mi_conv_bytes(int n, int h, int w, int cin, int cout, int k) {
    return 4.0 * ((double)n * h * w * (cin + cout) + (double)cout * k * k * cin);
}

static inline cudaStream_t mi_cs(mi_stream_t s) { return (cudaStream_t)s; }
static inline int mi_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline bool mi_al16(const void* p) { return (((uintptr_t)p) & 15u) == 0; }

// Round-to-nearest onto the TF32 grid (10-bit mantissa).  The tensor cores TRUNCATE fp32 operands to TF32, a bias of
// -2^-11 relative per operand that adds up coherently through a deep conv stack (measured: 0.056 dB of PSNR on the
// 256x448 SepConv task); values rounded here are read back exactly, so the only error left is unbiased.
__device__ __forceinline__ float mi_rn_tf32(float v) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
    return __uint_as_float(u);
}
__device__ __forceinline__ float4 mi_rn_tf32(float4 v) {
    return make_float4(mi_rn_tf32(v.x), mi_rn_tf32(v.y), mi_rn_tf32(v.z), mi_rn_tf32(v.w));
}
// host switch (MI_B200_TF32_RN=0 restores hardware truncation, for A/B measurements only)
bool mi_tf32_rn_enabled();

__device__ __forceinline__ float mi_act_apply(float v, int act, float slope) {
    switch (act) {
        case MI_ACT_RELU: return v > 0.f ? v : 0.f;
        case MI_ACT_LEAKY: return v > 0.f ? v : v * slope;
        case MI_ACT_SIGMOID: return 1.f / (1.f + __expf(-v));
        case MI_ACT_TANH: return tanhf(v);
        default: return v;
    }
}
// derivative expressed through the POST-activation value y
__device__ __forceinline__ float mi_act_grad(float y, int act, float slope) {
    switch (act) {
        case MI_ACT_RELU: return y > 0.f ? 1.f : 0.f;
        case MI_ACT_LEAKY: return y > 0.f ? 1.f : slope;
        case MI_ACT_SIGMOID: return y * (1.f - y);
        case MI_ACT_TANH: return 1.f - y * y;
        default: return 1.f;
    }
}

// shared by the SIMT and tcgen05 weight-gradient paths (conv_simt.cu)
int mi_bias_splits(long long m_total);
int mi_wgrad_finish_launch(const float* ws_w, const float* ws_b, int splits, int bias_splits, int cin, int cout, int k,
                           int ldw, int mode, float scale, float* grad_w, float* grad_b, const float* w_in, const float* b_in,
                           float* w_out, float* b_out, const float* lr_w, const float* lr_b, float* gsum_w,
                           float* gsum_b, float* wt_out, int ldwt, float* wr_out, cudaStream_t stream);
int mi_wgrad_splits(int n, int h, int wd, int cin, int cout, int k);
int mi_sm_budget();

// tcgen05 entry points (conv_tc.cu); return MI_ERR_UNSUPPORTED when the shape is not eligible
int mi_tc_fprop(const float* x, int ldx, const float* w, int ldw, const float* bias, float* y, int ldy,
                const float* mask_y, int ldmask, int mask_act, float mask_slope, int accumulate,
                int n, int h, int wd, int cin, int cout, int k, int act, float slope, cudaStream_t stream);
int mi_tc_wgrad_partials(const float* x, int ldx, const float* dy, int lddy, int n, int h, int wd, int cin, int cout,
                         int k, int ldw, float* ws_w, float* ws_b, int splits, int* bias_splits_out,
                         cudaStream_t stream);
bool mi_tc_fprop_eligible(const float* x, int ldx, const float* w, int ldw, const float* y, int ldy, int n, int h,
                          int wd, int cin, int cout, int k);
bool mi_tc_wgrad_eligible(const float* x, int ldx, const float* dy, int lddy, int n, int h, int wd, int cin, int cout,
                          int k);
bool mi_tc_wgrad_kx_shape(int cin, int cout, int k);
bool mi_tc_wgrad_kx_pair(int cin, int cout, int k);
