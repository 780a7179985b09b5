// tcgen05 (5th-gen tensor core) implicit-GEMM convolution for sm_100a: fprop (and dgrad through the
// rotated filter) and the split-K weight-gradient partials.  TF32 operands, fp32 accumulation in TMEM,
// operands staged by TMA with the 128-byte swizzle the UMMA shared-memory descriptors expect.
//
// fprop  : D[128 pixels][BN couts] += A[128 pixels][32 cin] * B[BN couts][32 cin]^T per (tap, cin chunk).
//          A is one TMA box {32 ch, TW, TH, 1} of the NHWC activation shifted by the tap; out-of-image
//          pixels and channel tails are zero-filled by TMA (that IS the conv padding).  Both operands K-major.
// wgrad  : D[128 couts][BN cins] += dY[P pixels][128 couts]^T * X_tap[P pixels][BN cins] per pixel chunk:
//          the reduction (pixels) is the slow dimension of both NHWC tensors, so both operands are MN-major.
//
// Warp roles per CTA: warp 0 = TMA producer (one lane), warp 1 = TMEM allocator + MMA issuer (one lane), then the
// epilogue warps (tcgen05.ld -> registers -> bias/activation/mask -> global): four in the per-tap kernels and the
// weight-gradient kernels (192 threads), eight in the halo kernels (320), eight plus a store warp in the default 3x3
// engine, the filter-column-stacked kernel of conv_tc_kxs.cuh (352).
// Replaces cuDNN's F.conv2d kernels used by the reference (model_utils.py:360).
#include <cuda.h>
#include <cstdio>
#include <cstdlib>

#include "mi_common.cuh"

namespace {

constexpr int BM = 128;          // UMMA M
constexpr int KCH = 32;          // fp32 channels per 128-byte swizzle row
constexpr int NTHREADS = 192;
constexpr uint32_t ROW_BYTES = 128;

// ------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                            int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// One lane of a fully converged warp.  Issuing tcgen05/TMA under `if (lane == 0)` makes the compiler treat the
// region as divergent and wrap every uniform-datapath instruction in an ELECT/BRA.U.ANY loop (seen in SASS, ~2x
// the issue cost per MMA); with elect.sync on warp-uniform control flow it emits straight-line code.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred P;\n"
        "elect.sync _|P, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, P;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout), 128-byte swizzle
// layout_type: 2 = SWIZZLE_128B (16-byte chunks; K-major operands), 1 = SWIZZLE_128B_BASE32B (32-byte chunks;
// the only layout the hardware accepts for MN-major TF32 operands)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint32_t layout_type = 2) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;                                  // descriptor version (Blackwell)
    d |= (uint64_t)((addr >> 7) & 7u) << 49;          // base offset: swizzle phase of the start row
    d |= (uint64_t)layout_type << 61;
    return d;
}
// The MMA-issuing thread is a single lane: building a descriptor from scratch per instruction (a dozen dependent
// integer ops) costs more issue cycles than a 128xN MMA with small N takes to execute.  So descriptors are built
// once per stage and advanced with one 64-bit add: the start-address field counts 16-byte units in bits [0,14).
__device__ __forceinline__ uint64_t desc_advance(uint64_t desc, uint32_t bytes) { return desc + (uint64_t)(bytes >> 4); }
// instruction descriptor for kind::tf32, fp32 accumulate (cute::UMMA::InstrDescriptor bit layout)
__device__ __forceinline__ uint32_t instr_desc(int m, int n, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// Epilogue for one 32-channel chunk of one output pixel (one thread == one TMEM lane).
// ncu showed the first version (per-element `if`s, __ldg of the bias inside them) serialising ~64 dependent
// global loads per tile; here the bias comes from shared memory, the activation switch is hoisted out of
// the element loop, and mask / accumulate operands are fetched as independent float4 loads.
struct EpiArgs {
    int cout, cout_store, act, mask_act, accumulate;   // cout_store = cout rounded up to 4 when the row stride allows
    float slope, mask_slope;
    int rnd;   // round what is stored to the TF32 grid (the next tensor-core conv then reads it exactly)
};

__device__ __forceinline__ void epilogue_chunk(const uint32_t (&v)[32], const float* __restrict__ sbias, int co,
                                               const EpiArgs& e, const float* __restrict__ mrow, float* yrow, bool vec) {
    float o[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(v[j]) + sbias[j];
    if (e.act == MI_ACT_RELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) o[j] = fmaxf(o[j], 0.f);
    } else if (e.act == MI_ACT_LEAKY) {
#pragma unroll
        for (int j = 0; j < 32; ++j) o[j] = o[j] > 0.f ? o[j] : o[j] * e.slope;
    } else if (e.act != MI_ACT_NONE) {
#pragma unroll
        for (int j = 0; j < 32; ++j) o[j] = mi_act_apply(o[j], e.act, e.slope);
    }
    if (vec && co + 32 <= e.cout) {
        if (mrow) {
            float4 m[8];
#pragma unroll
            for (int g = 0; g < 8; ++g) m[g] = __ldg(reinterpret_cast<const float4*>(mrow + co) + g);
            if (e.mask_act == MI_ACT_RELU) {
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    o[4 * g + 0] = m[g].x > 0.f ? o[4 * g + 0] : 0.f;
                    o[4 * g + 1] = m[g].y > 0.f ? o[4 * g + 1] : 0.f;
                    o[4 * g + 2] = m[g].z > 0.f ? o[4 * g + 2] : 0.f;
                    o[4 * g + 3] = m[g].w > 0.f ? o[4 * g + 3] : 0.f;
                }
            } else {
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    o[4 * g + 0] *= mi_act_grad(m[g].x, e.mask_act, e.mask_slope);
                    o[4 * g + 1] *= mi_act_grad(m[g].y, e.mask_act, e.mask_slope);
                    o[4 * g + 2] *= mi_act_grad(m[g].z, e.mask_act, e.mask_slope);
                    o[4 * g + 3] *= mi_act_grad(m[g].w, e.mask_act, e.mask_slope);
                }
            }
        }
        if (e.accumulate) {
            float4 a[8];
#pragma unroll
            for (int g = 0; g < 8; ++g) a[g] = *(reinterpret_cast<const float4*>(yrow + co) + g);
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                o[4 * g + 0] += a[g].x; o[4 * g + 1] += a[g].y; o[4 * g + 2] += a[g].z; o[4 * g + 3] += a[g].w;
            }
        }
        if (e.rnd) {
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = mi_rn_tf32(o[j]);
        }
#pragma unroll
        for (int g = 0; g < 8; ++g)
            *(reinterpret_cast<float4*>(yrow + co) + g) = make_float4(o[4 * g], o[4 * g + 1], o[4 * g + 2], o[4 * g + 3]);
    } else {
        // ragged tail (e.g. 51 channels): float4 groups while they fit in the padded row (the pad lane of a
        // 4-padded NHWC row is never read as data), guarded scalars for the rest
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            const int c = co + 4 * g;
            if (c >= e.cout) break;
            if (vec && c + 4 <= e.cout_store) {
                float4 r = make_float4(o[4 * g], o[4 * g + 1], o[4 * g + 2], o[4 * g + 3]);
                if (mrow) {
                    const float4 m = __ldg(reinterpret_cast<const float4*>(mrow + c));
                    r.x *= mi_act_grad(m.x, e.mask_act, e.mask_slope);
                    r.y *= mi_act_grad(m.y, e.mask_act, e.mask_slope);
                    r.z *= mi_act_grad(m.z, e.mask_act, e.mask_slope);
                    r.w *= mi_act_grad(m.w, e.mask_act, e.mask_slope);
                }
                if (e.accumulate) {
                    const float4 a = *reinterpret_cast<const float4*>(yrow + c);
                    r.x += a.x; r.y += a.y; r.z += a.z; r.w += a.w;
                }
                if (e.rnd) r = mi_rn_tf32(r);
                *reinterpret_cast<float4*>(yrow + c) = r;
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (c + q < e.cout) {
                        float val = o[4 * g + q];
                        if (mrow) val *= mi_act_grad(__ldg(mrow + c + q), e.mask_act, e.mask_slope);
                        if (e.accumulate) val += yrow[c + q];
                        yrow[c + q] = e.rnd ? mi_rn_tf32(val) : val;
                    }
                }
            }
        }
    }
}

// Warp-cooperative variant used whenever rows are 16-byte aligned.  After tcgen05.ld each lane holds 32 channels
// of ITS pixel, so a per-lane store touches 32 different 128-byte lines per instruction (ncu: the epilogue warps
// sat in the LSU queue for ~70% of a tile's period).  Here the 32x32 block is bounced through a padded per-warp
// shared-memory tile and read back transposed: 8 consecutive lanes then cover the 128 contiguous bytes of one
// pixel, i.e. 4 full lines per STG.128 instead of 32 partial ones; mask / accumulate operands are loaded with the
// same coalesced pattern.
struct RowMap {
    // The epilogue thread `lane` of TMEM-lane quarter q stores the staged rows i*4 + (lane>>3), i = 0..7.  Their global
    // pixel indices are the same for every 32-channel chunk of a tile, so they are resolved once per tile (the
    // per-chunk division by the tile width was a third of the epilogue's instructions).
    long long pix[8];
    uint32_t ok;          // bit i: row i lies inside the image
    __device__ __forceinline__ void init(int tw, int y0, int x0, int h, int w, int img, int q, int lane) {
        ok = 0u;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = q * 32 + i * 4 + (lane >> 3);
            const int th_i = r / tw, tw_i = r - th_i * tw;
            const int oy = y0 + th_i, ox = x0 + tw_i;
            pix[i] = ((long long)img * h + oy) * w + ox;
            if (oy < h && ox < w) ok |= 1u << i;
        }
    }
};
constexpr int EPI_PITCH = 36;   // floats per staged row (32 + 4): conflict-free for both access patterns

__device__ __forceinline__ void epilogue_chunk_coalesced(const uint32_t (&v)[32], const float* __restrict__ sbias,
                                                         int co, const EpiArgs& e, float* stage, int lane,
                                                         const RowMap& rm, const float* __restrict__ mask_y, int ldmask,
                                                         float* __restrict__ y, int ldy) {
    float o[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(v[j]) + sbias[j];
    if (e.act == MI_ACT_RELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) o[j] = fmaxf(o[j], 0.f);
    } else if (e.act == MI_ACT_LEAKY) {
#pragma unroll
        for (int j = 0; j < 32; ++j) o[j] = o[j] > 0.f ? o[j] : o[j] * e.slope;
    } else if (e.act != MI_ACT_NONE) {
#pragma unroll
        for (int j = 0; j < 32; ++j) o[j] = mi_act_apply(o[j], e.act, e.slope);
    }
    float* mine = stage + lane * EPI_PITCH;
#pragma unroll
    for (int g = 0; g < 8; ++g)
        *reinterpret_cast<float4*>(mine + 4 * g) = make_float4(o[4 * g], o[4 * g + 1], o[4 * g + 2], o[4 * g + 3]);
    __syncwarp();
    const int c4 = (lane & 7) * 4;          // channel offset of this lane's float4 inside the chunk
    const int c = co + c4;
    const bool full = c + 4 <= e.cout_store;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = i * 4 + (lane >> 3);
        const long long pix = rm.pix[i];
        if (!((rm.ok >> i) & 1u) || c >= e.cout) continue;
        float4 r = *reinterpret_cast<const float4*>(stage + row * EPI_PITCH + c4);
        float* dst = y + pix * ldy + c;
        if (full) {
            if (mask_y) {
                const float4 m = __ldg(reinterpret_cast<const float4*>(mask_y + pix * ldmask + c));
                r.x *= mi_act_grad(m.x, e.mask_act, e.mask_slope);
                r.y *= mi_act_grad(m.y, e.mask_act, e.mask_slope);
                r.z *= mi_act_grad(m.z, e.mask_act, e.mask_slope);
                r.w *= mi_act_grad(m.w, e.mask_act, e.mask_slope);
            }
            if (e.accumulate) {
                const float4 a = *reinterpret_cast<const float4*>(dst);
                r.x += a.x; r.y += a.y; r.z += a.z; r.w += a.w;
            }
            if (e.rnd) r = mi_rn_tf32(r);
            *reinterpret_cast<float4*>(dst) = r;
        } else {
            const float rv[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (c + q < e.cout) {
                    float val = rv[q];
                    if (mask_y) val *= mi_act_grad(__ldg(mask_y + pix * ldmask + c + q), e.mask_act, e.mask_slope);
                    if (e.accumulate) val += dst[q];
                    dst[q] = e.rnd ? mi_rn_tf32(val) : val;
                }
            }
        }
    }
    __syncwarp();
}

// 16-column pieces for the persistent halo kernels.  Per-role cycle counters showed their epilogue (one warp per
// TMEM lane quarter, ~1.8k cycles per 32x32 block of dependent shared/global round trips) pacing the MMA lane, so
// those kernels run EIGHT epilogue warps: the two warps that share a lane quarter each take one 16-column half of
// every 32-column chunk.  A staged row is 16 + 4 floats; 4 lanes cover the 64 contiguous bytes of one pixel.
constexpr int EPI16_PITCH = 20;
constexpr int HALO_THREADS = 320;   // warp 0 producer, warp 1 MMA, warps 2-9 epilogue

struct RowMap16 {
    long long pix[4];     // staged rows i*8 + (lane>>2), i = 0..3
    uint32_t ok;
    __device__ __forceinline__ void init(int tw, int y0, int x0, int h, int w, int img, int q, int lane) {
        ok = 0u;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = q * 32 + i * 8 + (lane >> 2);
            const int th_i = r / tw, tw_i = r - th_i * tw;
            const int oy = y0 + th_i, ox = x0 + tw_i;
            pix[i] = ((long long)img * h + oy) * w + ox;
            if (oy < h && ox < w) ok |= 1u << i;
        }
    }
};

__device__ __forceinline__ void epilogue_half_coalesced(const uint32_t (&v)[16], const float* __restrict__ sbias, int co,
                                                        const EpiArgs& e, float* stage, int lane, const RowMap16& rm,
                                                        const float* __restrict__ mask_y, int ldmask,
                                                        float* __restrict__ y, int ldy) {
    float o[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) o[j] = __uint_as_float(v[j]) + sbias[j];
    if (e.act == MI_ACT_RELU) {
#pragma unroll
        for (int j = 0; j < 16; ++j) o[j] = fmaxf(o[j], 0.f);
    } else if (e.act == MI_ACT_LEAKY) {
#pragma unroll
        for (int j = 0; j < 16; ++j) o[j] = o[j] > 0.f ? o[j] : o[j] * e.slope;
    } else if (e.act != MI_ACT_NONE) {
#pragma unroll
        for (int j = 0; j < 16; ++j) o[j] = mi_act_apply(o[j], e.act, e.slope);
    }
    float* mine = stage + lane * EPI16_PITCH;
#pragma unroll
    for (int g = 0; g < 4; ++g)
        *reinterpret_cast<float4*>(mine + 4 * g) = make_float4(o[4 * g], o[4 * g + 1], o[4 * g + 2], o[4 * g + 3]);
    __syncwarp();
    const int c4 = (lane & 3) * 4;
    const int c = co + c4;
    const bool full = c + 4 <= e.cout_store;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = i * 8 + (lane >> 2);
        const long long pix = rm.pix[i];
        if (!((rm.ok >> i) & 1u) || c >= e.cout) continue;
        float4 r = *reinterpret_cast<const float4*>(stage + row * EPI16_PITCH + c4);
        float* dst = y + pix * ldy + c;
        if (full) {
            if (mask_y) {
                const float4 m = __ldg(reinterpret_cast<const float4*>(mask_y + pix * ldmask + c));
                r.x *= mi_act_grad(m.x, e.mask_act, e.mask_slope);
                r.y *= mi_act_grad(m.y, e.mask_act, e.mask_slope);
                r.z *= mi_act_grad(m.z, e.mask_act, e.mask_slope);
                r.w *= mi_act_grad(m.w, e.mask_act, e.mask_slope);
            }
            if (e.accumulate) {
                const float4 a = *reinterpret_cast<const float4*>(dst);
                r.x += a.x; r.y += a.y; r.z += a.z; r.w += a.w;
            }
            if (e.rnd) r = mi_rn_tf32(r);
            *reinterpret_cast<float4*>(dst) = r;
        } else {
            const float rv[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (c + q < e.cout) {
                    float val = rv[q];
                    if (mask_y) val *= mi_act_grad(__ldg(mask_y + pix * ldmask + c + q), e.mask_act, e.mask_slope);
                    if (e.accumulate) val += dst[q];
                    dst[q] = e.rnd ? mi_rn_tf32(val) : val;
                }
            }
        }
    }
    __syncwarp();
}

// unaligned rows (channel slices of a concat buffer that do not start on a 16-byte boundary): plain per-lane stores
__device__ __forceinline__ void epilogue_half_scalar(const uint32_t (&v)[16], const float* __restrict__ sbias, int co,
                                                     const EpiArgs& e, const float* __restrict__ mrow, float* yrow) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        if (co + j >= e.cout) break;
        float val = mi_act_apply(__uint_as_float(v[j]) + sbias[j], e.act, e.slope);
        if (mrow) val *= mi_act_grad(__ldg(mrow + co + j), e.mask_act, e.mask_slope);
        if (e.accumulate) val += yrow[co + j];
        yrow[co + j] = e.rnd ? mi_rn_tf32(val) : val;
    }
}

struct FpropParams {
    int n, h, w, cin, cout, k, tw, th, tiles_x, tiles_y, bn, stages, act, accumulate, mask_act, ldy, ldmask, rnd;
    float slope, mask_slope;
    const float* bias;
    const float* mask_y;
    float* y;
};

// ------------------------------------------------------------------------------------------ fprop / dgrad
__global__ void __launch_bounds__(NTHREADS)
conv_fprop_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                     const FpropParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t a_bytes = BM * ROW_BYTES;
    const uint32_t b_bytes = (uint32_t)p.bn * ROW_BYTES;
    const uint32_t stage_bytes = a_bytes + b_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
    // bars[0..stages) = full, [stages..2*stages) = empty, [2*stages] = accumulator ready
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * p.stages + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int tile = blockIdx.x;
    const int tx_i = tile % p.tiles_x; tile /= p.tiles_x;
    const int ty_i = tile % p.tiles_y; tile /= p.tiles_y;
    const int img = tile;
    const int x0 = tx_i * p.tw, y0 = ty_i * p.th;
    const int co0 = blockIdx.y * p.bn;
    const int pad = p.k >> 1;
    const int chunks = (p.cin + KCH - 1) / KCH;
    const int iters = p.k * p.k * chunks;
    const int last_ksteps = (p.cin - (chunks - 1) * KCH + 7) / 8;

    __shared__ float sbias[256];
    // (the epilogue staging tile aliases pipeline stage 0: every MMA has retired before the epilogue starts)
    for (int i = threadIdx.x; i < p.bn; i += NTHREADS) {
        const int c = co0 + i;
        sbias[i] = (p.bias && c < p.cout) ? p.bias[c] : 0.f;
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(smem_u32(&bars[s]), 1);
            mbar_init(smem_u32(&bars[p.stages + s]), 1);
        }
        mbar_init(smem_u32(&bars[2 * p.stages]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), (uint32_t)p.bn);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // Non-persistent kernel with 2-3 co-resident CTAs per SM: only ONE lane of the producer / MMA warps spins on the
    // pipeline barriers (the other 31 park at the final __syncthreads); polling with the whole warp was measured
    // 15-40% slower on the deep layers because the pollers steal issue slots from the other CTAs' epilogues.
    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < iters; ++it) {
                const int s = it % p.stages;
                const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
                mbar_wait(smem_u32(&bars[p.stages + s]), ph ^ 1u);
                const int tap = it / chunks, ch = it - tap * chunks;
                const int ky = tap / p.k, kx = tap - ky * p.k;
                const uint32_t full = smem_u32(&bars[s]);
                const uint32_t a_dst = smem_u32(smem + (size_t)s * stage_bytes);
                mbar_expect_tx(full, stage_bytes);
                tma_load_4d(a_dst, &map_x, full, ch * KCH, x0 + kx - pad, y0 + ky - pad, img);
                tma_load_3d(a_dst + a_bytes, &map_w, full, ch * KCH, tap, co0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = instr_desc(BM, p.bn, 0, 0);
            for (int it = 0; it < iters; ++it) {
                const int s = it % p.stages;
                const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
                mbar_wait(smem_u32(&bars[s]), ph);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem + (size_t)s * stage_bytes);
                const uint32_t b_addr = a_addr + a_bytes;
                const uint64_t ad0 = smem_desc(a_addr, 16, 1024), bd0 = smem_desc(b_addr, 16, 1024);
                const bool last_chunk = (it % chunks) == chunks - 1;
                if (!last_chunk || last_ksteps == 4) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        umma_tf32(tmem_base, desc_advance(ad0, kk * 32), desc_advance(bd0, kk * 32), idesc,
                                  (it > 0 || kk > 0) ? 1u : 0u);
                } else {
                    // the last channel chunk holds fewer than 32 real channels: its all-zero K=8 steps are skipped
                    for (int kk = 0; kk < last_ksteps; ++kk)
                        umma_tf32(tmem_base, desc_advance(ad0, kk * 32), desc_advance(bd0, kk * 32), idesc,
                                  (it > 0 || kk > 0) ? 1u : 0u);
                }
                umma_commit(smem_u32(&bars[p.stages + s]));   // frees this smem stage when the MMAs retire
            }
            umma_commit(smem_u32(&bars[2 * p.stages]));       // accumulator complete
        }
    } else {
        // epilogue: warp q owns TMEM lanes [32q, 32q+32) == output pixels of the tile
        const int q = warp & 3;
        mbar_wait(smem_u32(&bars[2 * p.stages]), 0);
        tc_fence_after();
        const int r = q * 32 + lane;
        const int th_i = r / p.tw, tw_i = r - th_i * p.tw;
        const int oy = y0 + th_i, ox = x0 + tw_i;
        const bool pix_ok = (oy < p.h) && (ox < p.w);
        const long long pix = ((long long)img * p.h + oy) * p.w + ox;
        float* yrow = p.y + pix * p.ldy;
        const float* mrow = p.mask_y ? p.mask_y + pix * p.ldmask : nullptr;
        const bool vec = ((p.ldy & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.y) & 15) == 0) &&
                         (!p.mask_y || (((p.ldmask & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.mask_y) & 15) == 0)));
        EpiArgs ea;
        ea.cout = p.cout; ea.act = p.act; ea.mask_act = p.mask_act; ea.accumulate = p.accumulate; ea.rnd = p.rnd;
        ea.slope = p.slope; ea.mask_slope = p.mask_slope;
        ea.cout_store = p.cout;   // lanes past cout are never written: they may belong to the next concat slice
        RowMap rm;
        rm.init(p.tw, y0, x0, p.h, p.w, img, q, lane);
        float* stage = reinterpret_cast<float*>(smem) + q * 32 * EPI_PITCH;
        for (int c0 = 0; c0 < p.bn; c0 += 32) {
            if (co0 + c0 >= p.cout) break;             // warp-uniform
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            if (vec) epilogue_chunk_coalesced(v, sbias + c0, co0 + c0, ea, stage, lane, rm, p.mask_y, p.ldmask, p.y, p.ldy);
            else if (pix_ok) epilogue_chunk(v, sbias + c0, co0 + c0, ea, mrow, yrow, vec);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.bn);
}

// ------------------------------------------------------------------------------------------ fprop, small channels
// 3x3 layers with Cin <= 64 and Cout <= 64 (two thirds of SepConv's conv FLOPs, all at high resolution) are bound
// by L2->SM operand traffic in the per-tap kernel above (each tap re-reads the activation tile and every CTA
// re-reads the weights).  This persistent variant
//   * keeps ALL weights of the layer resident in shared memory (<= 144 KB), loaded once per CTA;
//   * stages one HALO tile of the activation per 32-channel chunk -- an 18x16-pixel box for an 8(w)x16(h)
//     output tile -- and issues the 9 taps' MMAs from that single copy by offsetting the UMMA descriptor
//     start address ((ky*16 + kx) rows of 128 B; the row pitch of 16 pixels keeps every 8-row core group at the
//     same swizzle phase kx, carried in the descriptor's base-offset field);
//   * double-buffers the accumulator in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.
// The halo row pitch is 10 pixels (8 + 2): because the swizzle follows absolute shared-memory address bits, the
// 8-row core groups may sit at any 16-byte-aligned offset, so the stride between them (SBO) is simply the pitch
// (1280 B) and no padding rows are staged.  MI_B200_HALO_PITCH=16 restores the padded 16-pixel pitch (SBO 2048).
constexpr int HALO_H = 18, HT_W = 8, HT_H = 16;

struct HaloParams {
    int rnd;
    int n, h, w, cin, cout, chunks, bn, stages, tiles_x, tiles_y, total_tiles, act, accumulate, mask_act, ldy, ldmask,
        halo_w;
    uint32_t halo_bytes, halo_stride;   // bytes one TMA box delivers / 1024-aligned distance between stages
    unsigned long long* dbg;            // optional per-role cycle counters of CTA 0 (MI_B200_DEBUG_TIMING=1)
    float slope, mask_slope;
    const float* bias;
    const float* mask_y;
    float* y;
};

__global__ void __launch_bounds__(HALO_THREADS)
conv_fprop_tc_halo_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                          const HaloParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t b_tile = (uint32_t)p.bn * ROW_BYTES;            // one (tap, chunk) weight tile
    const uint32_t b_total = 9u * p.chunks * b_tile;
    uint8_t* smem_a = smem + b_total;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_a + (size_t)p.stages * p.halo_stride);
    // bars: [0,S) full, [S,2S) empty, 2S = weights of chunk 0 loaded, 2S+1..2 = tmem full[2], 2S+3..4 = tmem empty[2],
    // 2S+5 = weights of chunk 1 loaded
    const int S = p.stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 6);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    __shared__ float sbias[64];
    __shared__ __align__(16) float epi_stage[8 * 32 * EPI16_PITCH];
    if (threadIdx.x < 64) sbias[threadIdx.x] = (p.bias && (int)threadIdx.x < p.cout) ? p.bias[threadIdx.x] : 0.f;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(smem_u32(&bars[s]), 1);
            mbar_init(smem_u32(&bars[S + s]), 1);
        }
        mbar_init(smem_u32(&bars[2 * S]), 1);
        mbar_init(smem_u32(&bars[2 * S + 1]), 1);
        mbar_init(smem_u32(&bars[2 * S + 2]), 1);
        mbar_init(smem_u32(&bars[2 * S + 3]), 256);     // all eight epilogue warps release an accumulator buffer
        mbar_init(smem_u32(&bars[2 * S + 4]), 256);
        mbar_init(smem_u32(&bars[2 * S + 5]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), (uint32_t)(2 * p.bn));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int my_tiles = (p.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (warp == 0) {
        // The resident weights arrive per 32-channel chunk, each behind its own barrier, with the first halo box in
        // between: the MMAs of chunk 0 start as soon as its nine weight tiles (half of the 147 KB of a 64-channel
        // layer) and the first halo are in, instead of after the whole filter bank (the wait was ~4.5 us per launch)
        if (elect_one()) {
            const uint32_t wbar = smem_u32(&bars[2 * S]);
            mbar_expect_tx(wbar, 9u * b_tile);
            for (int tap = 0; tap < 9; ++tap)
                tma_load_3d(smem_u32(smem) + (uint32_t)(tap * p.chunks) * b_tile, &map_w, wbar, 0, tap, 0);
        }
        __syncwarp();
        int it = 0;
        long long t_empty = 0;
        const long long t_begin = clock64();
        for (int t = 0; t < my_tiles; ++t) {
            int tile = (int)blockIdx.x + t * (int)gridDim.x;
            const int tx_i = tile % p.tiles_x; tile /= p.tiles_x;
            const int ty_i = tile % p.tiles_y; tile /= p.tiles_y;
            const int img = tile;
            for (int ch = 0; ch < p.chunks; ++ch, ++it) {
                const int s = it % S;
                const uint32_t ph = (uint32_t)(it / S) & 1u;
                const long long c0 = clock64();
                mbar_wait(smem_u32(&bars[S + s]), ph ^ 1u);
                t_empty += clock64() - c0;
                if (elect_one()) {
                    const uint32_t full = smem_u32(&bars[s]);
                    mbar_expect_tx(full, p.halo_bytes);
                    tma_load_4d(smem_u32(smem_a + (size_t)s * p.halo_stride), &map_x, full, ch * KCH, tx_i * HT_W - 1,
                                ty_i * HT_H - 1, img);
                    if (it == 0 && p.chunks > 1) {
                        const uint32_t wbar = smem_u32(&bars[2 * S + 5]);
                        mbar_expect_tx(wbar, 9u * b_tile);
                        for (int tap = 0; tap < 9; ++tap)
                            tma_load_3d(smem_u32(smem) + (uint32_t)(tap * p.chunks + 1) * b_tile, &map_w, wbar, KCH, tap, 0);
                    }
                }
                __syncwarp();
            }
        }
        if (my_tiles == 0 && p.chunks > 1 && elect_one()) {
            // (a CTA without tiles still owes the second barrier its bytes before the block may retire)
            const uint32_t wbar = smem_u32(&bars[2 * S + 5]);
            mbar_expect_tx(wbar, 9u * b_tile);
            for (int tap = 0; tap < 9; ++tap)
                tma_load_3d(smem_u32(smem) + (uint32_t)(tap * p.chunks + 1) * b_tile, &map_w, wbar, KCH, tap, 0);
        }
        if (p.dbg && blockIdx.x == 0 && lane == 0) {
            p.dbg[0] = (unsigned long long)t_empty; p.dbg[1] = (unsigned long long)(clock64() - t_begin);
        }
    } else if (warp == 1) {
        const uint32_t idesc = instr_desc(BM, p.bn, 0, 0);
        const long long t_begin = clock64();
        long long t_w = 0, t_full = 0, t_tmem = 0;
        mbar_wait(smem_u32(&bars[2 * S]), 0);
        t_w = clock64() - t_begin;
        tc_fence_after();
        if (my_tiles == 0 && p.chunks > 1) mbar_wait(smem_u32(&bars[2 * S + 5]), 0);   // in-flight TMA must land
        const uint32_t b_base = smem_u32(smem);
        const int last_ksteps = (p.cin - (p.chunks - 1) * KCH + 7) / 8;
        int it = 0;
        for (int t = 0; t < my_tiles; ++t) {
            const int buf = t & 1;
            const uint32_t use = (uint32_t)(t >> 1);
            long long c0 = clock64();
            mbar_wait(smem_u32(&bars[2 * S + 3 + buf]), (use & 1u) ^ 1u);   // epilogue drained this buffer
            t_tmem += clock64() - c0;
            tc_fence_after();
            const uint32_t d_addr = tmem_base + (uint32_t)(buf * p.bn);
            for (int ch = 0; ch < p.chunks; ++ch, ++it) {
                const int s = it % S;
                const uint32_t ph = (uint32_t)(it / S) & 1u;
                c0 = clock64();
                mbar_wait(smem_u32(&bars[s]), ph);
                t_full += clock64() - c0;
                if (t == 0 && ch == 1) {
                    c0 = clock64();
                    mbar_wait(smem_u32(&bars[2 * S + 5]), 0);      // second half of the filter bank
                    t_w += clock64() - c0;
                }
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a_base = smem_u32(smem_a + (size_t)s * p.halo_stride);
                    const uint64_t ad0 = smem_desc(a_base, 16, (uint32_t)p.halo_w * ROW_BYTES);   // 1024-aligned base
                    const uint64_t bd0 = smem_desc(b_base + (uint32_t)ch * b_tile, 16, 1024);
                    const uint32_t tap_b = (uint32_t)p.chunks * b_tile;
                    // Views shifted by (ky halo rows + kx pixels).  Measured on B200 (tools/diag_halo.py): the 128B
                    // swizzle XOR is taken from the ABSOLUTE smem address bits [7,10), so a row-shifted start keeps
                    // base offset 0; putting the start row's phase there (as the PTX text suggests) corrupts kx != 0.
                    if (ch < p.chunks - 1 || last_ksteps == 4) {
#pragma unroll
                        for (int tap = 0; tap < 9; ++tap) {
                            const int ky = tap / 3, kx = tap - ky * 3;
                            const uint64_t a_tap = desc_advance(ad0, (uint32_t)(ky * p.halo_w + kx) * ROW_BYTES);
                            const uint64_t b_tap = desc_advance(bd0, (uint32_t)tap * tap_b);
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk)
                                umma_tf32(d_addr, desc_advance(a_tap, kk * 32), desc_advance(b_tap, kk * 32), idesc,
                                          (ch > 0 || tap > 0 || kk > 0) ? 1u : 0u);
                        }
                    } else {
                        // ragged last chunk (e.g. 51 = 32 + 19 channels): skip its all-zero K=8 steps.  Unrolled per
                        // step count: the rolled form cost ~100 cycles per MMA against ~57 for the unrolled one
                        // (per-role cycle counters), i.e. the single issuing lane, not the tensor pipe, set the pace
#define MI_HALO_RAGGED(KS)                                                                                          \
    _Pragma("unroll") for (int tap = 0; tap < 9; ++tap) {                                                           \
        const int ky = tap / 3, kx = tap - ky * 3;                                                                  \
        const uint64_t a_tap = desc_advance(ad0, (uint32_t)(ky * p.halo_w + kx) * ROW_BYTES);                       \
        const uint64_t b_tap = desc_advance(bd0, (uint32_t)tap * tap_b);                                            \
        _Pragma("unroll") for (int kk = 0; kk < KS; ++kk)                                                           \
            umma_tf32(d_addr, desc_advance(a_tap, kk * 32), desc_advance(b_tap, kk * 32), idesc,                    \
                      (ch > 0 || tap > 0 || kk > 0) ? 1u : 0u);                                                     \
    }
                        if (last_ksteps == 3) { MI_HALO_RAGGED(3) }
                        else if (last_ksteps == 2) { MI_HALO_RAGGED(2) }
                        else { MI_HALO_RAGGED(1) }
                    }
                    umma_commit(smem_u32(&bars[S + s]));
                    if (ch == p.chunks - 1) umma_commit(smem_u32(&bars[2 * S + 1 + buf]));
                }
                __syncwarp();
            }
        }
        if (p.dbg && blockIdx.x == 0 && lane == 0) {
            p.dbg[2] = (unsigned long long)t_w; p.dbg[3] = (unsigned long long)t_full;
            p.dbg[4] = (unsigned long long)t_tmem; p.dbg[5] = (unsigned long long)(clock64() - t_begin);
        }
    } else {
        const int q = warp & 3;               // TMEM lane quarter this warp may read
        const int half = (warp - 2) >> 2;     // which 16-column half of every 32-column chunk it owns
        const int r = q * 32 + lane;
        const int th_i = r >> 3, tw_i = r & 7;
        long long e_wait = 0, e_ld = 0, e_st = 0;
        const long long e_begin = clock64();
        const bool vec = ((p.ldy & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.y) & 15) == 0) &&
                         (!p.mask_y || (((p.ldmask & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.mask_y) & 15) == 0)));
        EpiArgs ea;
        ea.cout = p.cout; ea.act = p.act; ea.mask_act = p.mask_act; ea.accumulate = p.accumulate; ea.rnd = p.rnd;
        ea.slope = p.slope; ea.mask_slope = p.mask_slope;
        ea.cout_store = p.cout;   // lanes past cout are never written: they may belong to the next concat slice
        float* stage = epi_stage + (warp - 2) * 32 * EPI16_PITCH;
        for (int t = 0; t < my_tiles; ++t) {
            const int buf = t & 1;
            const uint32_t use = (uint32_t)(t >> 1);
            int tile = (int)blockIdx.x + t * (int)gridDim.x;
            const int tx_i = tile % p.tiles_x; tile /= p.tiles_x;
            const int ty_i = tile % p.tiles_y; tile /= p.tiles_y;
            const int img = tile;
            const int oy = ty_i * HT_H + th_i, ox = tx_i * HT_W + tw_i;
            const bool pix_ok = (oy < p.h) && (ox < p.w);
            const long long pix = ((long long)img * p.h + oy) * p.w + ox;
            float* yrow = p.y + pix * p.ldy;
            const float* mrow = p.mask_y ? p.mask_y + pix * p.ldmask : nullptr;
            RowMap16 rm;
            rm.init(HT_W, ty_i * HT_H, tx_i * HT_W, p.h, p.w, img, q, lane);
            long long c0 = clock64();
            mbar_wait(smem_u32(&bars[2 * S + 1 + buf]), use & 1u);
            e_wait += clock64() - c0;
            tc_fence_after();
            for (int c0 = 16 * half; c0 < p.bn; c0 += 32) {
                if (c0 >= p.cout) break;            // warp-uniform
                uint32_t v[16];
                long long c1 = clock64();
                tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * p.bn + c0), v);
                e_ld += clock64() - c1;
                c1 = clock64();
                if (vec) epilogue_half_coalesced(v, sbias + c0, c0, ea, stage, lane, rm, p.mask_y, p.ldmask, p.y, p.ldy);
                else if (pix_ok) epilogue_half_scalar(v, sbias + c0, c0, ea, mrow, yrow);
                e_st += clock64() - c1;
            }
            tc_fence_before();
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars[2 * S + 3 + buf])) : "memory");
        }
        if (p.dbg && blockIdx.x == 0 && threadIdx.x == 64) {
            p.dbg[6] = (unsigned long long)e_wait; p.dbg[7] = (unsigned long long)e_ld;
            p.dbg[8] = (unsigned long long)e_st; p.dbg[9] = (unsigned long long)(clock64() - e_begin);
            p.dbg[10] = (unsigned long long)my_tiles;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)(2 * p.bn));
}

// ------------------------------------------------------------------------------------------ fprop, mid/deep channels
// 3x3 layers with more than 64 channels: the weights no longer fit in shared memory next to the halo ring, so they
// are streamed -- but at the SAME granularity as the halo tile: one pipeline stage = the 18x10-pixel halo box of a
// 32-channel chunk PLUS the nine {64 couts x 32 cin} weight tiles of that chunk, behind ONE mbarrier.  Measured on
// B200 (per-role cycle counters): the MMA lane pays ~450 cycles of fixed cost per barrier wait (try_wait, fence,
// elect, descriptor set-up, commit), so the per-tap kernel (4 MMAs per wait) and a per-tap weight ring (4-8 MMAs per
// wait) both leave the tensor pipe idle 70 % of the time, while 36 MMAs per wait amortise it (this is why the
// small-channel halo kernel is the fastest of the family).  Persistent CTAs walk (pixel tile, 64-cout tile) work
// items; accumulators are double buffered in TMEM so the epilogue of one item overlaps the MMAs of the next.
struct HaloStreamParams {
    int rnd;
    int n, h, w, cin, cout, chunks, n_tiles, tiles_x, tiles_y, total_tiles, items, act, accumulate, mask_act, ldy,
        ldmask, halo_w;
    uint32_t halo_bytes, halo_stride;
    unsigned long long* dbg;            // optional per-role cycle counters of CTA 0 (MI_B200_DEBUG_TIMING=1)
    float slope, mask_slope;
    const float* bias;
    const float* mask_y;
    float* y;
};
constexpr int HS_BN = 64;        // cout tile of the streamed kernel
constexpr int HS_STAGES = 2;

__global__ void __launch_bounds__(HALO_THREADS)
conv_fprop_tc_halo_stream_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                                 const HaloStreamParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int S = HS_STAGES;
    constexpr uint32_t b_tile = HS_BN * ROW_BYTES;                 // one (tap, chunk) weight tile: 8 KB
    const uint32_t stage_bytes = p.halo_stride + 9u * b_tile;      // halo box + nine weight tiles
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)S * stage_bytes);
    // bars: [0,S) full, [S,2S) empty, 2S..2S+1 tmem full[2], 2S+2..2S+3 tmem empty[2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long t_start = clock64();

    __shared__ float sbias[512];
    __shared__ __align__(16) float epi_stage[8 * 32 * EPI16_PITCH];
    for (int i = threadIdx.x; i < 512; i += HALO_THREADS) sbias[i] = (p.bias && i < p.cout) ? p.bias[i] : 0.f;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(smem_u32(&bars[s]), 1);
            mbar_init(smem_u32(&bars[S + s]), 1);
        }
        mbar_init(smem_u32(&bars[2 * S]), 1);
        mbar_init(smem_u32(&bars[2 * S + 1]), 1);
        mbar_init(smem_u32(&bars[2 * S + 2]), 256);     // all eight epilogue warps release an accumulator buffer
        mbar_init(smem_u32(&bars[2 * S + 3]), 256);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), (uint32_t)(2 * HS_BN));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int my_items = (p.items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (warp == 0) {
        int it = 0;
        long long w_e = 0;
        for (int t = 0; t < my_items; ++t) {
            const int item = (int)blockIdx.x + t * (int)gridDim.x;
            const int nt = item % p.n_tiles;
            int tile = item / p.n_tiles;
            const int tx_i = tile % p.tiles_x; tile /= p.tiles_x;
            const int ty_i = tile % p.tiles_y; tile /= p.tiles_y;
            const int img = tile;
            for (int ch = 0; ch < p.chunks; ++ch, ++it) {
                const int s = it % S;
                const long long c0 = clock64();
                mbar_wait(smem_u32(&bars[S + s]), (((uint32_t)(it / S)) & 1u) ^ 1u);
                w_e += clock64() - c0;
                if (elect_one()) {
                    const uint32_t full = smem_u32(&bars[s]);
                    const uint32_t base = smem_u32(smem + (size_t)s * stage_bytes);
                    mbar_expect_tx(full, p.halo_bytes + 9u * b_tile);
                    tma_load_4d(base, &map_x, full, ch * KCH, tx_i * HT_W - 1, ty_i * HT_H - 1, img);
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap)
                        tma_load_3d(base + p.halo_stride + (uint32_t)tap * b_tile, &map_w, full, ch * KCH, tap,
                                    nt * HS_BN);
                }
                __syncwarp();
            }
        }
        if (p.dbg && blockIdx.x == 0 && lane == 0) {
            p.dbg[0] = (unsigned long long)w_e; p.dbg[2] = (unsigned long long)(clock64() - t_start);
        }
    } else if (warp == 1) {
        const uint32_t idesc = instr_desc(BM, HS_BN, 0, 0);
        const int last_ksteps = (p.cin - (p.chunks - 1) * KCH + 7) / 8;
        int it = 0;
        long long w_f = 0, w_te = 0, t_first = 0;
        for (int t = 0; t < my_items; ++t) {
            const int buf = t & 1;
            long long c0 = clock64();
            mbar_wait(smem_u32(&bars[2 * S + 2 + buf]), (((uint32_t)(t >> 1)) & 1u) ^ 1u);   // epilogue drained it
            w_te += clock64() - c0;
            tc_fence_after();
            const uint32_t d_addr = tmem_base + (uint32_t)(buf * HS_BN);
            for (int ch = 0; ch < p.chunks; ++ch, ++it) {
                const int s = it % S;
                c0 = clock64();
                mbar_wait(smem_u32(&bars[s]), ((uint32_t)(it / S)) & 1u);
                w_f += clock64() - c0;
                if (it == 0) t_first = clock64() - t_start;
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a_base = smem_u32(smem + (size_t)s * stage_bytes);
                    const uint64_t ad0 = smem_desc(a_base, 16, (uint32_t)p.halo_w * ROW_BYTES);
                    const uint64_t bd0 = smem_desc(a_base + p.halo_stride, 16, 1024);
                    if (ch < p.chunks - 1 || last_ksteps == 4) {
#pragma unroll
                        for (int tap = 0; tap < 9; ++tap) {
                            const int ky = tap / 3, kx = tap - ky * 3;
                            const uint64_t a_tap = desc_advance(ad0, (uint32_t)(ky * p.halo_w + kx) * ROW_BYTES);
                            const uint64_t b_tap = desc_advance(bd0, (uint32_t)tap * b_tile);
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk)
                                umma_tf32(d_addr, desc_advance(a_tap, kk * 32), desc_advance(b_tap, kk * 32), idesc,
                                          (ch > 0 || tap > 0 || kk > 0) ? 1u : 0u);
                        }
                    } else {
                        const uint32_t tap_b = b_tile;
                        if (last_ksteps == 3) { MI_HALO_RAGGED(3) }
                        else if (last_ksteps == 2) { MI_HALO_RAGGED(2) }
                        else { MI_HALO_RAGGED(1) }
                    }
                    umma_commit(smem_u32(&bars[S + s]));
                    if (ch == p.chunks - 1) umma_commit(smem_u32(&bars[2 * S + buf]));
                }
                __syncwarp();
            }
        }
        if (p.dbg && blockIdx.x == 0 && lane == 0) {
            p.dbg[3] = (unsigned long long)w_f; p.dbg[5] = (unsigned long long)w_te;
            p.dbg[6] = (unsigned long long)t_first; p.dbg[7] = (unsigned long long)(clock64() - t_start);
        }
    } else {
        const int q = warp & 3;               // TMEM lane quarter this warp may read
        const int half = (warp - 2) >> 2;     // which 16-column half of every 32-column chunk it owns
        const bool vec = ((p.ldy & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.y) & 15) == 0) &&
                         (!p.mask_y || (((p.ldmask & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.mask_y) & 15) == 0)));
        EpiArgs ea;
        ea.cout = p.cout; ea.act = p.act; ea.mask_act = p.mask_act; ea.accumulate = p.accumulate; ea.rnd = p.rnd;
        ea.slope = p.slope; ea.mask_slope = p.mask_slope;
        ea.cout_store = p.cout;   // lanes past cout are never written: they may belong to the next concat slice
        const int r = q * 32 + lane;
        const int th_i = r >> 3, tw_i = r & 7;
        float* stage = epi_stage + (warp - 2) * 32 * EPI16_PITCH;
        long long e_wait = 0;
        for (int t = 0; t < my_items; ++t) {
            const int buf = t & 1;
            const int item = (int)blockIdx.x + t * (int)gridDim.x;
            const int nt = item % p.n_tiles;
            const int co0 = nt * HS_BN;
            int tile = item / p.n_tiles;
            const int tx_i = tile % p.tiles_x; tile /= p.tiles_x;
            const int ty_i = tile % p.tiles_y; tile /= p.tiles_y;
            const int img = tile;
            const int oy = ty_i * HT_H + th_i, ox = tx_i * HT_W + tw_i;
            const bool pix_ok = (oy < p.h) && (ox < p.w);
            const long long pix = ((long long)img * p.h + oy) * p.w + ox;
            float* yrow = p.y + pix * p.ldy;
            const float* mrow = p.mask_y ? p.mask_y + pix * p.ldmask : nullptr;
            const long long c0 = clock64();
            mbar_wait(smem_u32(&bars[2 * S + buf]), ((uint32_t)(t >> 1)) & 1u);
            e_wait += clock64() - c0;
            tc_fence_after();
            RowMap16 rm;
            rm.init(HT_W, ty_i * HT_H, tx_i * HT_W, p.h, p.w, img, q, lane);
#pragma unroll
            for (int c0i = 16 * half; c0i < HS_BN; c0i += 32) {
                if (co0 + c0i >= p.cout) break;                    // warp-uniform
                uint32_t v[16];
                tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * HS_BN + c0i), v);
                if (vec) epilogue_half_coalesced(v, sbias + co0 + c0i, co0 + c0i, ea, stage, lane, rm, p.mask_y, p.ldmask,
                                                 p.y, p.ldy);
                else if (pix_ok) epilogue_half_scalar(v, sbias + co0 + c0i, co0 + c0i, ea, mrow, yrow);
            }
            tc_fence_before();
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars[2 * S + 2 + buf])) : "memory");
        }
        if (p.dbg && blockIdx.x == 0 && threadIdx.x == 64) {
            p.dbg[8] = (unsigned long long)e_wait; p.dbg[9] = (unsigned long long)(clock64() - t_start);
            p.dbg[10] = (unsigned long long)my_items;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)(2 * HS_BN));
}

#include "conv_tc_kxs.cuh"

// ------------------------------------------------------------------------------------------ fprop, 5x5 / 7x7
// The 5x5 / 7x7 layers of superslomo and voxelflow.  Their filter bank is too large for a pipeline stage (49 taps x 32
// couts x 128 B = 196 KB per 32-channel chunk), so it streams one FILTER ROW at a time: the halo box of a chunk
// ((16+k-1) x (8+k-1) pixels, one TMA box, ring of two) stays put while k rows of k weight tiles pass through their
// own ring (three stages); the MMA lane waits once per row, i.e. once per 4k = 20-28 MMAs.  Tap (ky, kx) is again a
// descriptor view of the halo box shifted by ky rows and kx pixels.  Persistent CTAs over (pixel tile, cout tile) items,
// double-buffered TMEM accumulators, the eight-warp epilogue of the 3x3 halo kernels.
struct HaloRowsParams {
    int rnd;
    int n, h, w, cin, cout, k, chunks, bn, n_tiles, tiles_x, tiles_y, total_tiles, items, act, accumulate, mask_act,
        ldy, ldmask, halo_w, sa, sb;
    uint32_t halo_bytes, halo_stride;
    float slope, mask_slope;
    const float* bias;
    const float* mask_y;
    float* y;
};

__global__ void __launch_bounds__(HALO_THREADS)
conv_fprop_tc_halo_rows_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                               const HaloRowsParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int SA = p.sa, SB = p.sb, K = p.k, pad = p.k >> 1;
    const uint32_t b_tile = (uint32_t)p.bn * ROW_BYTES;            // one (tap, chunk) weight tile
    const uint32_t b_stage = (uint32_t)K * b_tile;                 // one filter row of a chunk
    uint8_t* smem_b = smem + (size_t)SA * p.halo_stride;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + (size_t)SB * b_stage);
    uint64_t* a_full = bars;
    uint64_t* a_empty = bars + SA;
    uint64_t* b_full = bars + 2 * SA;
    uint64_t* b_empty = bars + 2 * SA + SB;
    uint64_t* t_full = bars + 2 * SA + 2 * SB;
    uint64_t* t_empty = t_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t tmem_cols = (uint32_t)(2 * p.bn);               // 64 or 128

    __shared__ float sbias[512];
    __shared__ __align__(16) float epi_stage[8 * 32 * EPI16_PITCH];
    for (int i = threadIdx.x; i < 512; i += HALO_THREADS) sbias[i] = (p.bias && i < p.cout) ? p.bias[i] : 0.f;
    if (threadIdx.x == 0) {
        for (int s = 0; s < SA; ++s) { mbar_init(smem_u32(&a_full[s]), 1); mbar_init(smem_u32(&a_empty[s]), 1); }
        for (int s = 0; s < SB; ++s) { mbar_init(smem_u32(&b_full[s]), 1); mbar_init(smem_u32(&b_empty[s]), 1); }
        mbar_init(smem_u32(&t_full[0]), 1);
        mbar_init(smem_u32(&t_full[1]), 1);
        mbar_init(smem_u32(&t_empty[0]), 256);
        mbar_init(smem_u32(&t_empty[1]), 256);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int my_items = (p.items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (warp == 0) {
        int ia = 0, ib = 0;
        for (int t = 0; t < my_items; ++t) {
            const int item = (int)blockIdx.x + t * (int)gridDim.x;
            const int nt = item % p.n_tiles;
            int tile = item / p.n_tiles;
            const int tx_i = tile % p.tiles_x; tile /= p.tiles_x;
            const int ty_i = tile % p.tiles_y; tile /= p.tiles_y;
            const int img = tile;
            for (int ch = 0; ch < p.chunks; ++ch, ++ia) {
                const int s = ia % SA;
                mbar_wait(smem_u32(&a_empty[s]), (((uint32_t)(ia / SA)) & 1u) ^ 1u);
                if (elect_one()) {
                    const uint32_t full = smem_u32(&a_full[s]);
                    mbar_expect_tx(full, p.halo_bytes);
                    tma_load_4d(smem_u32(smem + (size_t)s * p.halo_stride), &map_x, full, ch * KCH, tx_i * HT_W - pad,
                                ty_i * HT_H - pad, img);
                }
                __syncwarp();
                for (int ky = 0; ky < K; ++ky, ++ib) {
                    const int sb = ib % SB;
                    mbar_wait(smem_u32(&b_empty[sb]), (((uint32_t)(ib / SB)) & 1u) ^ 1u);
                    if (elect_one()) {
                        const uint32_t full = smem_u32(&b_full[sb]);
                        mbar_expect_tx(full, b_stage);
                        for (int kx = 0; kx < K; ++kx)
                            tma_load_3d(smem_u32(smem_b + (size_t)sb * b_stage) + (uint32_t)kx * b_tile, &map_w, full,
                                        ch * KCH, ky * K + kx, nt * p.bn);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = instr_desc(BM, p.bn, 0, 0);
        const int last_ksteps = (p.cin - (p.chunks - 1) * KCH + 7) / 8;
        int ia = 0, ib = 0;
        for (int t = 0; t < my_items; ++t) {
            const int buf = t & 1;
            mbar_wait(smem_u32(&t_empty[buf]), (((uint32_t)(t >> 1)) & 1u) ^ 1u);   // epilogue drained this buffer
            tc_fence_after();
            const uint32_t d_addr = tmem_base + (uint32_t)(buf * p.bn);
            for (int ch = 0; ch < p.chunks; ++ch, ++ia) {
                const int s = ia % SA;
                mbar_wait(smem_u32(&a_full[s]), ((uint32_t)(ia / SA)) & 1u);
                tc_fence_after();
                const uint32_t a_base = smem_u32(smem + (size_t)s * p.halo_stride);
                const int ksteps = (ch == p.chunks - 1) ? last_ksteps : 4;
                for (int ky = 0; ky < K; ++ky, ++ib) {
                    const int sb = ib % SB;
                    mbar_wait(smem_u32(&b_full[sb]), ((uint32_t)(ib / SB)) & 1u);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t ad0 = smem_desc(a_base, 16, (uint32_t)p.halo_w * ROW_BYTES);
                        const uint64_t bd0 = smem_desc(smem_u32(smem_b + (size_t)sb * b_stage), 16, 1024);
                        for (int kx = 0; kx < K; ++kx) {
                            const uint64_t a_tap = desc_advance(ad0, (uint32_t)(ky * p.halo_w + kx) * ROW_BYTES);
                            const uint64_t b_tap = desc_advance(bd0, (uint32_t)kx * b_tile);
                            if (ksteps == 4) {
#pragma unroll
                                for (int kk = 0; kk < 4; ++kk)
                                    umma_tf32(d_addr, desc_advance(a_tap, kk * 32), desc_advance(b_tap, kk * 32), idesc,
                                              (ch > 0 || ky > 0 || kx > 0 || kk > 0) ? 1u : 0u);
                            } else {
                                for (int kk = 0; kk < ksteps; ++kk)
                                    umma_tf32(d_addr, desc_advance(a_tap, kk * 32), desc_advance(b_tap, kk * 32), idesc,
                                              (ch > 0 || ky > 0 || kx > 0 || kk > 0) ? 1u : 0u);
                            }
                        }
                        umma_commit(smem_u32(&b_empty[sb]));
                        if (ky == K - 1) {
                            umma_commit(smem_u32(&a_empty[s]));
                            if (ch == p.chunks - 1) umma_commit(smem_u32(&t_full[buf]));
                        }
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        const int q = warp & 3;               // TMEM lane quarter this warp may read
        const int half = (warp - 2) >> 2;     // which 16-column half of every 32-column chunk it owns
        const bool vec = ((p.ldy & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.y) & 15) == 0) &&
                         (!p.mask_y || (((p.ldmask & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.mask_y) & 15) == 0)));
        EpiArgs ea;
        ea.cout = p.cout; ea.act = p.act; ea.mask_act = p.mask_act; ea.accumulate = p.accumulate; ea.rnd = p.rnd;
        ea.slope = p.slope; ea.mask_slope = p.mask_slope;
        ea.cout_store = p.cout;
        const int r = q * 32 + lane;
        const int th_i = r >> 3, tw_i = r & 7;
        float* stage = epi_stage + (warp - 2) * 32 * EPI16_PITCH;
        for (int t = 0; t < my_items; ++t) {
            const int buf = t & 1;
            const int item = (int)blockIdx.x + t * (int)gridDim.x;
            const int nt = item % p.n_tiles;
            const int co0 = nt * p.bn;
            int tile = item / p.n_tiles;
            const int tx_i = tile % p.tiles_x; tile /= p.tiles_x;
            const int ty_i = tile % p.tiles_y; tile /= p.tiles_y;
            const int img = tile;
            const int oy = ty_i * HT_H + th_i, ox = tx_i * HT_W + tw_i;
            const bool pix_ok = (oy < p.h) && (ox < p.w);
            const long long pix = ((long long)img * p.h + oy) * p.w + ox;
            float* yrow = p.y + pix * p.ldy;
            const float* mrow = p.mask_y ? p.mask_y + pix * p.ldmask : nullptr;
            mbar_wait(smem_u32(&t_full[buf]), ((uint32_t)(t >> 1)) & 1u);
            tc_fence_after();
            RowMap16 rm;
            rm.init(HT_W, ty_i * HT_H, tx_i * HT_W, p.h, p.w, img, q, lane);
            for (int c0i = 16 * half; c0i < p.bn; c0i += 32) {
                if (co0 + c0i >= p.cout) break;                    // warp-uniform
                uint32_t v[16];
                tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * p.bn + c0i), v);
                if (vec) epilogue_half_coalesced(v, sbias + co0 + c0i, co0 + c0i, ea, stage, lane, rm, p.mask_y, p.ldmask,
                                                 p.y, p.ldy);
                else if (pix_ok) epilogue_half_scalar(v, sbias + co0 + c0i, co0 + c0i, ea, mrow, yrow);
            }
            tc_fence_before();
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&t_empty[buf])) : "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

// ------------------------------------------------------------------------------------------ wgrad partials
struct WgradParams {
    int n, h, w, cin, cout, k, ldw, pw, ph, tiles_x, tiles_y, bn, stages, co_tiles, ci_tiles, tiles_per_split,
        total_tiles;
    float* ws_w;
};

__global__ void __launch_bounds__(NTHREADS)
conv_wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x,
                     const WgradParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int P = p.pw * p.ph;                          // pixels per stage (multiple of 8)
    const uint32_t box_bytes = (uint32_t)P * ROW_BYTES;  // one {32 ch x P pixels} box
    const int a_boxes = BM / KCH;                        // 4 boxes of 32 couts
    const int b_boxes = p.bn / KCH;
    const uint32_t a_bytes = a_boxes * box_bytes, b_bytes = b_boxes * box_bytes;
    const uint32_t stage_bytes = a_bytes + b_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * p.stages + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int bid = blockIdx.x;
    const int ci_tile = bid % p.ci_tiles; bid /= p.ci_tiles;
    const int co_tile = bid % p.co_tiles; bid /= p.co_tiles;
    const int tap = bid;
    const int ky = tap / p.k, kx = tap - ky * p.k, pad = p.k >> 1;
    const int split = blockIdx.y;
    const int t_begin = split * p.tiles_per_split;
    const int t_end = min(t_begin + p.tiles_per_split, p.total_tiles);
    const int iters = max(t_end - t_begin, 0);
    const int co0 = co_tile * BM, ci0 = ci_tile * p.bn;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(smem_u32(&bars[s]), 1);
            mbar_init(smem_u32(&bars[p.stages + s]), 1);
        }
        mbar_init(smem_u32(&bars[2 * p.stages]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), (uint32_t)p.bn);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < iters; ++it) {
                const int s = it % p.stages;
                const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
                mbar_wait(smem_u32(&bars[p.stages + s]), ph ^ 1u);
                int t = t_begin + it;
                const int tx_i = t % p.tiles_x; t /= p.tiles_x;
                const int ty_i = t % p.tiles_y; t /= p.tiles_y;
                const int img = t;
                const int x0 = tx_i * p.pw, y0 = ty_i * p.ph;
                const uint32_t full = smem_u32(&bars[s]);
                const uint32_t base = smem_u32(smem + (size_t)s * stage_bytes);
                mbar_expect_tx(full, stage_bytes);
                for (int j = 0; j < a_boxes; ++j)
                    tma_load_4d(base + j * box_bytes, &map_dy, full, co0 + j * KCH, x0, y0, img);
                for (int j = 0; j < b_boxes; ++j)
                    tma_load_4d(base + a_bytes + j * box_bytes, &map_x, full, ci0 + j * KCH, x0 + kx - pad,
                                y0 + ky - pad, img);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = instr_desc(BM, p.bn, 1, 1);   // both operands MN-major
            const int ksteps = P / 8;
            for (int it = 0; it < iters; ++it) {
                const int s = it % p.stages;
                const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
                mbar_wait(smem_u32(&bars[s]), ph);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem + (size_t)s * stage_bytes);
                const uint32_t b_addr = a_addr + a_bytes;
                // MN-major TF32: swizzle atoms are 4 pixels (K) x 32 channels (MN) = 512 B, so one K=8 MMA
                // spans two atoms SBO=512 B apart; MN blocks of 32 channels are box_bytes apart (LBO)
                const uint64_t ad0 = smem_desc(a_addr, box_bytes, 512, 1), bd0 = smem_desc(b_addr, box_bytes, 512, 1);
                for (int kk = 0; kk < ksteps; ++kk)
                    umma_tf32(tmem_base, desc_advance(ad0, kk * 1024), desc_advance(bd0, kk * 1024), idesc,
                              (it > 0 || kk > 0) ? 1u : 0u);
                umma_commit(smem_u32(&bars[p.stages + s]));
            }
            umma_commit(smem_u32(&bars[2 * p.stages]));
        }
    } else {
        const int q = warp & 3;
        const int co = co0 + q * 32 + lane;
        float* dst = p.ws_w + (long long)split * p.cout * p.k * p.k * p.ldw + ((long long)co * p.k * p.k + tap) * p.ldw;
        if (iters > 0) {
            mbar_wait(smem_u32(&bars[2 * p.stages]), 0);
            tc_fence_after();
        }
        for (int c0 = 0; c0 < p.bn; c0 += 32) {
            uint32_t v[32];
            if (iters > 0) {
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0u;
            }
            if (co >= p.cout) continue;
            const int cb = ci0 + c0;
            if (cb >= p.cin) continue;
            if (cb + 32 <= p.ldw) {
                // whole chunk inside the padded row: 8 float4 stores (columns >= cin are exact zeros: TMA zero-fill)
#pragma unroll
                for (int g = 0; g < 8; ++g)
                    *(reinterpret_cast<float4*>(dst + cb) + g) =
                        make_float4(__uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1]),
                                    __uint_as_float(v[4 * g + 2]), __uint_as_float(v[4 * g + 3]));
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (cb + j < p.cin) dst[cb + j] = __uint_as_float(v[j]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.bn);
}

// ------------------------------------------------------------------------------------------ wgrad, small channels
// The per-tap kernel above re-reads dY and X from L2 once per tap: 18 activation passes per layer, which is what
// bounds it on the full-resolution 32/64-channel layers (ncu: ~10 TB/s of L2->SM traffic, M=128 MMAs with 32-64 real
// rows).  For 3x3 layers with Cin <= 64 and Cout <= 128 a CTA owns one filter COLUMN kx and all three rows ky:
//   * per 8(w) x R(h) pixel tile it loads dY once and ONE X box of R+2 rows shifted by kx-1 columns;
//   * the three ky taps are the same box seen from a start address ky tile-rows (ky * 1024 B) further down -- a whole
//     number of swizzle periods, so no descriptor trickery is involved -- and they are stacked along the MMA N
//     dimension by setting the descriptor's MN-block stride (LBO) to one tile row: one 128 x 96 x 8 MMA per
//     32-channel block and tile row computes D[cout][(ky, cin)] for all three ky at once.
// L2 traffic drops from 18 to ~6.4 activation passes and the MMA count per pixel by 3x.
struct WgradKxParams {
    int n, h, w, cin, cout, k, ldw, rows, tiles_x, tiles_y, total_tiles, tiles_per_split, stages, na, nb, ci_tiles;
    float* ws_w;
    float* ws_b;      // != nullptr: the bias-gradient partial of every split is produced here too (see the epilogue warps)
    int merged;       // 3x3, Cin a multiple of 64: both 32-channel halves of the X block in ONE N = 192 MMA per tile row
    int pair;         // Cout <= 64: TWO filter columns per CTA in the two halves of the M = 128 accumulator (see kernel)
};

__global__ void __launch_bounds__(NTHREADS)
conv_wgrad_tc_kx_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x,
                        const WgradKxParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t row_bytes = 8u * ROW_BYTES;                       // one tile row: 8 pixels x 32 channels
    const uint32_t a_box = (uint32_t)p.rows * row_bytes;             // dY box {32 ch, 8 px, R rows}
    const uint32_t b_box = (uint32_t)(p.rows + p.k - 1) * row_bytes; // X box  {32 ch, 8 px, R+k-1 rows}
    const int pad = p.k >> 1;
    const uint32_t ncols = 32u * (uint32_t)p.k;                      // accumulator columns per X box: k taps x 32 cin
    // Paired mode (Cout <= 64, i.e. at most two dY boxes): with one filter column per CTA half of the 128 accumulator
    // rows were padding.  The gradient of tap (ky, kx) is sum_q X[q + ky - pad][ci] * dY[q - (kx - pad)][co] over the X
    // pixels q of the tile, so ONE unshifted X box serves two filter columns if the dY boxes are the ones that shift:
    // A = [dY shifted for kx_a | dY shifted for kx_b] fills all four MN blocks, rows 0..63 of D belong to kx_a and
    // rows 64..127 (32..63 for Cout <= 32) to kx_b, and a layer needs ceil(k / 2) CTAs per split-K slice instead of k.
    const int a_slots = p.pair ? 2 * p.na : p.na;
    const uint32_t a_bytes = (uint32_t)a_slots * a_box, b_bytes = (uint32_t)p.nb * b_box;
    const uint32_t stage_bytes = a_bytes + b_bytes;
    // the M=128 A descriptor always spans four 32-channel blocks; blocks past `na` alias whatever follows (the X
    // boxes, the next stage, or the zeroed tail pad after the last stage) and only feed accumulator rows >= cout,
    // which are never read back
    const uint32_t tail_pad = (uint32_t)(4 - a_slots) * a_box;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes + tail_pad);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * p.stages + 1);
    uint32_t tmem_cols = 32;                                         // nb * k * 32 accumulator columns -> power of two
    while (tmem_cols < (uint32_t)p.nb * ncols) tmem_cols <<= 1;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // blockIdx.x = kx + 3 * (cin tile of 64 + ci_tiles * cout tile of 128): wider layers are cut into channel blocks,
    // each block re-reading its dY / X boxes (3 * ci_tiles and 3 * co_tiles passes instead of 9 * tiles of the per-tap form)
    const int groups = p.pair ? (p.k + 1) / 2 : p.k;                // filter-column groups per (cin, cout) block
    const int kx = p.pair ? 2 * ((int)blockIdx.x % groups) : (int)blockIdx.x % groups;   // (first) filter column of this CTA
    const bool kxb_ok = p.pair && kx + 1 < p.k;                    // paired mode: the second column exists
    const int ct = (int)blockIdx.x / groups;
    const int ci0 = (ct % p.ci_tiles) * 64, co0 = (ct / p.ci_tiles) * BM;
    const int split = blockIdx.y;
    const int t_begin = split * p.tiles_per_split;
    const int t_end = min(t_begin + p.tiles_per_split, p.total_tiles);
    const int iters = max(t_end - t_begin, 0);

    // finite contents for the aliased A blocks of the last stage
    for (uint32_t i = threadIdx.x; i < tail_pad / 16; i += NTHREADS)
        reinterpret_cast<float4*>(smem + (size_t)p.stages * stage_bytes)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    // The bias gradient rides along: the dY boxes of every tile pass through this CTA's shared memory anyway, and the
    // four epilogue warps have nothing to do until the accumulator is final.  In the CTA of the centre filter column
    // and the first cin block, epilogue warp q sums box q over the pixels of every stage (lane l = channel
    // co0 + 32 q + l) and releases the stage together with the MMA lane (the `empty` barrier then counts 1 + na
    // arrivals); one partial per split-K slice replaces a separate pass over dY (`bias_partial_kernel`) per layer.
    // (paired mode: the unshifted dY, the one of the centre column, is slot 0 or slot 1 of this CTA's pair)
    const int bias_slot = pad - kx;
    const bool do_bias = p.ws_b != nullptr && ci0 == 0 && (p.pair ? (bias_slot == 0 || (bias_slot == 1 && kxb_ok)) : kx == pad);
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(smem_u32(&bars[s]), 1);
            mbar_init(smem_u32(&bars[p.stages + s]), do_bias ? 1u + (uint32_t)p.na : 1u);
        }
        mbar_init(smem_u32(&bars[2 * p.stages]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), tmem_cols);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // tail-pad stores visible to the MMA's async proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // Both single-lane roles run on warp-uniform control flow with elect.sync: under `if (lane == 0)` the compiler wraps
    // every TMA / tcgen05 instruction in an ELECT + BRA.U.ANY loop (seen in the SASS of the first version of this
    // kernel), which with the rolled row loop made the issuing lane, not the tensor pipe, set the pace.
    if (warp == 0) {
        int s = 0;
        uint32_t ph = 0;
        for (int it = 0; it < iters; ++it) {
            mbar_wait(smem_u32(&bars[p.stages + s]), ph ^ 1u);
            if (elect_one()) {
                int t = t_begin + it;
                const int tx_i = t % p.tiles_x; t /= p.tiles_x;
                const int ty_i = t % p.tiles_y; t /= p.tiles_y;
                const int img = t;
                const int x0 = tx_i * 8, y0 = ty_i * p.rows;
                const uint32_t full = smem_u32(&bars[s]);
                const uint32_t base = smem_u32(smem + (size_t)s * stage_bytes);
                // paired mode: X stays put and the dY boxes shift (slot 0: column kx, slot 1: column kx + 1); otherwise
                // dY stays put and X shifts
                const int xs = p.pair ? 0 : kx - pad;
                mbar_expect_tx(full, stage_bytes - ((p.pair && !kxb_ok) ? (uint32_t)p.na * a_box : 0u));
                for (int j = 0; j < p.na; ++j)
                    tma_load_4d(base + j * a_box, &map_dy, full, co0 + j * KCH, x0 + (p.pair ? pad - kx : 0), y0, img);
                if (kxb_ok)
                    for (int j = 0; j < p.na; ++j)
                        tma_load_4d(base + (p.na + j) * a_box, &map_dy, full, co0 + j * KCH, x0 + pad - kx - 1, y0, img);
                if (p.merged) {
                    // one 5-D box {32 ch, 8 px, 2 halves, R+2 rows}: in shared memory the two 32-channel halves of a tile
                    // row follow each other, so (ky, half) are six MN blocks at ONE stride of 1024 bytes
                    tma_load_5d(base + a_bytes, &map_x, full, 0, x0 + xs, ci0 / KCH, y0 - pad, img);
                } else {
                    for (int j = 0; j < p.nb; ++j)
                        tma_load_4d(base + a_bytes + j * b_box, &map_x, full, ci0 + j * KCH, x0 + xs, y0 - pad, img);
                }
            }
            __syncwarp();
            if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
    } else if (warp == 1) {
        const uint32_t idesc = instr_desc(BM, (int)ncols, 1, 1);   // both operands MN-major, N = k taps x 32 channels
        const uint32_t idesc2 = instr_desc(BM, 2 * (int)ncols, 1, 1);
        // A: 4 MN blocks (32 couts each) one dY box apart; B: k MN blocks (ky = 0..k-1) one tile row apart.
        // K = 8 pixels = one tile row = two 4-pixel swizzle atoms 512 B apart (SBO).
        const uint64_t ad_base = smem_desc(smem_u32(smem), a_box, 512, 1);
        const uint64_t bd_base = smem_desc(smem_u32(smem) + a_bytes, row_bytes, 512, 1);
        int s = 0;
        uint32_t ph = 0, s_off = 0;
        for (int it = 0; it < iters; ++it) {
            mbar_wait(smem_u32(&bars[s]), ph);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t ad0 = desc_advance(ad_base, s_off);
                if (p.merged) {
                    // N = 192: the dY fetch of a tile row is shared by six (ky, half) blocks instead of three
                    const uint64_t bd0 = desc_advance(bd_base, s_off);
#define MI_WGKX_ROWS2(R0, R1)                                                                                       \
    _Pragma("unroll") for (int r = R0; r < R1; ++r)                                                                 \
        umma_tf32(tmem_base, desc_advance(ad0, (uint32_t)r * 1024u), desc_advance(bd0, (uint32_t)r * 2048u), idesc2,\
                  (it > 0 || r > 0) ? 1u : 0u);
                    MI_WGKX_ROWS2(0, 8)
                    if (p.rows == 16) { MI_WGKX_ROWS2(8, 16) }
#undef MI_WGKX_ROWS2
                } else
                for (int j = 0; j < p.nb; ++j) {
                    const uint64_t bd0 = desc_advance(bd_base, s_off + (uint32_t)j * b_box);
                    const uint32_t d_addr = tmem_base + (uint32_t)j * ncols;
#define MI_WGKX_ROWS(R0, R1)                                                                                        \
    _Pragma("unroll") for (int r = R0; r < R1; ++r)                                                                 \
        umma_tf32(d_addr, desc_advance(ad0, (uint32_t)r * 1024u), desc_advance(bd0, (uint32_t)r * 1024u), idesc,   \
                  (it > 0 || r > 0) ? 1u : 0u);
                    MI_WGKX_ROWS(0, 8)
                    if (p.rows == 16) { MI_WGKX_ROWS(8, 16) }
#undef MI_WGKX_ROWS
                }
                umma_commit(smem_u32(&bars[p.stages + s]));
                if (it + 1 == iters) umma_commit(smem_u32(&bars[2 * p.stages]));
            }
            __syncwarp();
            if (++s == p.stages) { s = 0; s_off = 0; ph ^= 1u; } else { s_off += stage_bytes; }
        }
    } else {
        const int q = warp & 3;
        // accumulator rows 32 q .. 32 q + 31: cout block q -- or, in paired mode, cout block q % na of filter column
        // kx + q / na
        const int slot = p.pair ? q / p.na : 0;
        const int kx_q = kx + slot;
        const bool rows_ok = p.pair ? (slot == 0 || (slot == 1 && kxb_ok)) : true;
        const int co = co0 + (p.pair ? q % p.na : q) * 32 + lane;
        const int kk2 = p.k * p.k;
        if (do_bias && (p.pair ? (slot == bias_slot) : q < p.na)) {
            // box rows are 128-byte pixel rows in the SWIZZLE_128B_ATOM_32B pattern (cute's Swizzle<2,5,2>): the
            // 32-byte chunk index is XOR-ed with the pixel-row index mod 4
            float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            const int npix = p.rows * 8;
            const uint32_t lane_off = ((uint32_t)(lane & 7)) << 2, chunk = (uint32_t)lane >> 3;
            int s = 0;
            uint32_t ph = 0;
            for (int it = 0; it < iters; ++it) {
                mbar_wait(smem_u32(&bars[s]), ph);
                const uint32_t bj = smem_u32(smem + (size_t)s * stage_bytes) + (uint32_t)q * a_box + lane_off;
                for (int px = 0; px < npix; px += 8) {          // pixel row px + i has swizzle phase i & 3
                    float v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        asm volatile("ld.shared.f32 %0, [%1];"
                                     : "=f"(v[i]) : "r"(bj + (uint32_t)(px + i) * ROW_BYTES + ((chunk ^ (uint32_t)(i & 3)) << 5)));
#pragma unroll
                    for (int i = 0; i < 8; ++i) a[i] += v[i];
                }
                __syncwarp();
                if (lane == 0)
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars[p.stages + s])) : "memory");
                if (++s == p.stages) { s = 0; ph ^= 1u; }
            }
            if (co < p.cout)
                p.ws_b[(long long)split * p.cout + co] = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
        }
        float* dst0 = p.ws_w + (long long)split * p.cout * kk2 * p.ldw + (long long)co * kk2 * p.ldw;
        if (iters > 0) {
            mbar_wait(smem_u32(&bars[2 * p.stages]), 0);
            tc_fence_after();
        }
        for (int j = 0; j < p.nb; ++j) {
            for (int ky = 0; ky < p.k; ++ky) {
                uint32_t v[32];
                if (iters > 0) {
                    const uint32_t col = p.merged ? (uint32_t)((ky * 2 + j) * 32) : (uint32_t)j * ncols + (uint32_t)(ky * 32);
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + col, v);
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = 0u;
                }
                if (co >= p.cout || !rows_ok) continue;
                const int cb = ci0 + j * KCH;
                if (cb >= p.cin) continue;
                float* dst = dst0 + (ky * p.k + kx_q) * p.ldw;
                if (cb + 32 <= p.ldw) {
#pragma unroll
                    for (int g = 0; g < 8; ++g)
                        *(reinterpret_cast<float4*>(dst + cb) + g) =
                            make_float4(__uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1]),
                                        __uint_as_float(v[4 * g + 2]), __uint_as_float(v[4 * g + 3]));
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (cb + i < p.cin) dst[cb + i] = __uint_as_float(v[i]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

// column sums of dy for the bias gradient.  grid = (pixel slices, 32-channel groups); a block is 8 float4 channel
// lanes x 32 pixel lanes with 4 independent loads in flight per thread (16 KB per block), so a few hundred blocks
// keep enough bytes in flight to stream dy near HBM rate; one deterministic partial per (slice, channel).
__global__ void __launch_bounds__(256)
bias_partial_kernel(const float* __restrict__ dy, int lddy, float* __restrict__ ws_b, int cout, long long m_total,
                    long long chunk, int vec) {
    const int split = blockIdx.x;
    const long long m0 = (long long)split * chunk;
    long long m1 = m0 + chunk;
    if (m1 > m_total) m1 = m_total;
    __shared__ float red[32][33];
    const int lane_c = threadIdx.x & 7, lane_p = threadIdx.x >> 3;
    const int c = blockIdx.y * 32 + lane_c * 4;
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    if (c < cout) {
        if (vec) {
            float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0, s2 = s0, s3 = s0;
            long long m = m0 + lane_p;
            for (; m + 96 < m1; m += 128) {
                const float4 v0 = *reinterpret_cast<const float4*>(dy + m * lddy + c);
                const float4 v1 = *reinterpret_cast<const float4*>(dy + (m + 32) * lddy + c);
                const float4 v2 = *reinterpret_cast<const float4*>(dy + (m + 64) * lddy + c);
                const float4 v3 = *reinterpret_cast<const float4*>(dy + (m + 96) * lddy + c);
                s0.x += v0.x; s0.y += v0.y; s0.z += v0.z; s0.w += v0.w;
                s1.x += v1.x; s1.y += v1.y; s1.z += v1.z; s1.w += v1.w;
                s2.x += v2.x; s2.y += v2.y; s2.z += v2.z; s2.w += v2.w;
                s3.x += v3.x; s3.y += v3.y; s3.z += v3.z; s3.w += v3.w;
            }
            for (; m < m1; m += 32) {
                const float4 v0 = *reinterpret_cast<const float4*>(dy + m * lddy + c);
                s0.x += v0.x; s0.y += v0.y; s0.z += v0.z; s0.w += v0.w;
            }
            a[0] = (s0.x + s1.x) + (s2.x + s3.x); a[1] = (s0.y + s1.y) + (s2.y + s3.y);
            a[2] = (s0.z + s1.z) + (s2.z + s3.z); a[3] = (s0.w + s1.w) + (s2.w + s3.w);
        } else {
            for (long long m = m0 + lane_p; m < m1; m += 32)
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (c + q < cout) a[q] += dy[m * lddy + c + q];
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) red[lane_p][lane_c * 4 + q] = a[q];
    __syncthreads();
    if (threadIdx.x < 32) {
        const int cc = blockIdx.y * 32 + threadIdx.x;
        if (cc < cout) {
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) s += red[j][threadIdx.x];
            ws_b[(long long)split * cout + cc] = s;
        }
    }
}

// ------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}

// NHWC activation view [n][h][w][c] (pixel stride ld floats) with box {32, bw, bh, 1}
bool make_act_map(CUtensorMap* map, const float* base, int ld, int n, int h, int w, int c, int bw, int bh,
                  CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
    cuuint64_t strides[3] = {(cuuint64_t)ld * 4, (cuuint64_t)w * ld * 4, (cuuint64_t)h * w * ld * 4};
    cuuint32_t box[4] = {(cuuint32_t)KCH, (cuuint32_t)bw, (cuuint32_t)bh, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// KRSC weights [cout][k*k][cin] (row stride ldw floats) with box {32, 1, bn}
bool make_weight_map(CUtensorMap* map, const float* base, int ldw, int cout, int taps, int cin, int bn) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[3] = {(cuuint64_t)cin, (cuuint64_t)taps, (cuuint64_t)cout};
    cuuint64_t strides[2] = {(cuuint64_t)ldw * 4, (cuuint64_t)taps * ldw * 4};
    cuuint32_t box[3] = {(cuuint32_t)KCH, 1, (cuuint32_t)bn};
    cuuint32_t estr[3] = {1, 1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int device_is_sm100() {
    static int cached = -1;
    if (cached < 0) {
        int dev = 0, major = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
            cudaGetLastError();
            return 0;
        }
        cached = (major == 10) ? 1 : 0;
    }
    return cached;
}

// spatial tile of `pixels` pixels: widest power-of-two row segment that fits the image width
void pick_tile(int w, int pixels, int* tw, int* th) {
    int t = 1;
    while (t * 2 <= pixels && t * 2 <= w) t *= 2;
    // prefer an exact divisor of w when one exists among the powers of two (no wasted columns)
    int best = t;
    for (int c = t; c >= 8; c >>= 1)
        if (w % c == 0) { best = c; break; }
    *tw = best;
    *th = pixels / best;
}

int pick_bn(int cout) {
    if (cout <= 32) return 32;
    if (cout <= 64) return 64;
    if (cout <= 128) return 128;
    return (cout % 256 == 0 || cout > 384) ? 256 : 128;
}

bool aligned_view(const float* p, int ld) { return mi_al16(p) && (ld % 4 == 0); }

// MI_B200_HALO=0 falls back to the per-tap kernel for the small-channel layers (A/B switch for profiling)
bool halo_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("MI_B200_HALO");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

// MI_B200_HALO_STREAM=0 keeps the per-tap kernel for the >64-channel 3x3 layers (A/B switch for profiling)
bool halo_stream_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("MI_B200_HALO_STREAM");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

int halo_pitch() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("MI_B200_HALO_PITCH");
        v = (e && atoi(e) == 16) ? 16 : 10;
    }
    return v;
}

int num_sms() { return mi_sm_budget(); }   // CTAs a persistent launch may occupy (conv_simt.cu)

// Filter-column-stacked 3x3 kernel (conv_tc_kxs.cuh).  MI_B200_KXS=0 keeps the halo kernels (A/B switch), =2 forces the
// stacked kernel even where its 14-of-16-column tiles cover the image worse than the 8x16 halo tiles, =3 additionally
// forces its generic (per-thread store) epilogue, =4 keeps the two-box staging tile everywhere (A/B of the third stage).
int kxs_mode() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("MI_B200_KXS");
        v = e ? atoi(e) : 1;
    }
    return v;
}
bool kxs_enabled(int n, int h, int wd) {
    // Measured on B200 (tools/bench_conv.py, profiles/r02_conv_kxs_vs_halo.txt): faster than the halo kernels on every
    // SepConv layer shape, including the 12x16 / 24x32 ones where 14-column tiles cover the image badly (more, shorter
    // work items there), so the stacked kernel takes every eligible 3x3 layer.
    (void)n; (void)h; (void)wd;
    return kxs_mode() != 0;
}

int g_pad_lanes_scratch = 0;

int launch_kxs(const float* x, int ldx, const float* w, int ldw, const float* bias, float* y, int ldy,
               const float* mask_y, int ldmask, int mask_act, float mask_slope, int accumulate, int n, int h, int wd,
               int cin, int cout, int act, float slope, cudaStream_t stream) {
    const bool stream_w = cin > 64 || cout > 64;
    KxsParams kp;
    kp.n = n; kp.h = h; kp.w = wd; kp.cin = cin; kp.cout = cout;
    kp.chunks = mi_cdiv(cin, KCH);
    kp.bn = stream_w ? 64 : (cout <= 32 ? 32 : 64);
    kp.tiles_x = mi_cdiv(wd, KX_OW);
    kp.tiles_y = mi_cdiv(h, KX_H);
    kp.total_tiles = kp.tiles_x * kp.tiles_y * n;
    kp.n_tiles = stream_w ? mi_cdiv(cout, kp.bn) : 1;
    kp.items = kp.total_tiles * kp.n_tiles;
    kp.act = act; kp.slope = slope; kp.accumulate = accumulate; kp.mask_act = mask_act; kp.rnd = mi_tf32_rn_enabled();
    kp.mask_slope = mask_slope; kp.ldy = ldy; kp.ldmask = ldmask; kp.bias = bias; kp.mask_y = mask_y; kp.y = y;
    const size_t b_chunk = (size_t)9 * kp.bn * ROW_BYTES;
    // dynamic shared memory: [resident weights] [S stages] [epilogue staging 28 KB] [barriers] + 1 KB alignment slack;
    // 227 KB per CTA minus the static part (bias table)
    const size_t budget = 227 * 1024 - 3072 - 128;   // static: bias table, tmem slot, alignment
    const bool vec = aligned_view(y, ldy) && (!mask_y || aligned_view(mask_y, ldmask));
    int epi = KXS_EPI_GENERIC;
    if (vec && kxs_mode() != 3) {
        if (!mask_y && !accumulate) epi = KXS_EPI_PLAIN;
        else if (!(mask_y && accumulate)) epi = KXS_EPI_OPERAND;
    }
    size_t fixed = KX_STAGING + 256 + 1024;
    size_t smem;
    kp.one_box = 0;
    if (stream_w) {
        kp.stages = 2;
        smem = 2 * (KX_BOX_BYTES + b_chunk) + fixed;
        if (smem > budget) return MI_ERR_UNSUPPORTED;
    } else {
        const size_t b_total = kp.chunks * b_chunk;
        int stages = (int)((budget - fixed - b_total) / KX_BOX_BYTES);
        // two stages leave the box loads exposed (per-role counters: the MMA lane waits ~330 cycles per stage for data
        // on the 64-channel layers, whose 147 KB of weights leave room for no more); a plain epilogue can go through
        // one 14 KB staging box in two rounds and give the room to a third stage
        if (stages == 2 && epi == KXS_EPI_PLAIN && kxs_mode() == 5 &&
            (budget - (fixed - KX_OUT_BOX) - b_total) / KX_BOX_BYTES >= 3) {
            kp.one_box = 1;
            fixed -= KX_OUT_BOX;
            stages = 3;
        }
        if (stages > 4) stages = 4;
        if (stages < 2) return MI_ERR_UNSUPPORTED;
        kp.stages = stages;
        smem = b_total + (size_t)stages * KX_BOX_BYTES + fixed;
    }
    CUtensorMap map_x, map_w, map_y, map_op;
    if (!make_act_map(&map_x, x, ldx, n, h, wd, cin, KX_W, KX_BOX_H)) return MI_ERR_UNSUPPORTED;
    if (!make_weight_map(&map_w, w, ldw, cout, 9, cin, kp.bn)) return MI_ERR_UNSUPPORTED;
    if (epi != KXS_EPI_GENERIC) {
        // channels below cout & ~3 only (see the ragged-tail note in conv_tc_kxs.cuh) -- unless the pad lanes of the
        // row are the caller's to overwrite (mi_set_pad_lanes_scratch), then the whole padded row goes through TMA
        const int cpad = (cout + 3) & ~3;
        const bool pad_ok = g_pad_lanes_scratch && ldy == cpad && (!mask_y || ldmask == cpad);
        kp.c_tma = pad_ok ? cpad : (cout & ~3);
        if (!make_act_map(&map_y, y, ldy, n, h, wd, kp.c_tma, KX_OW, KX_H)) return MI_ERR_UNSUPPORTED;
        if (epi == KXS_EPI_OPERAND && mask_y) {
            if (!make_act_map(&map_op, mask_y, ldmask, n, h, wd, kp.c_tma, KX_OW, KX_H)) return MI_ERR_UNSUPPORTED;
        } else {
            map_op = map_y;
        }
    } else {
        map_y = map_x; map_op = map_x;     // unused
        kp.c_tma = cout & ~3;
    }
    typedef void (*KxsKernel)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const KxsParams);
    static const KxsKernel kernels[2][3] = {
        {conv_fprop_tc_kxs_kernel<false, 0>, conv_fprop_tc_kxs_kernel<false, 1>, conv_fprop_tc_kxs_kernel<false, 2>},
        {conv_fprop_tc_kxs_kernel<true, 0>, conv_fprop_tc_kxs_kernel<true, 1>, conv_fprop_tc_kxs_kernel<true, 2>}};
    static bool attr = false;
    if (!attr) {
        for (int a = 0; a < 2; ++a)
            for (int b = 0; b < 3; ++b) {
                cudaError_t e = cudaFuncSetAttribute(kernels[a][b], cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                     (int)budget);
                if (e != cudaSuccess) return (int)e;
            }
        attr = true;
    }
    const int grid = kp.items < num_sms() ? kp.items : num_sms();
    static unsigned long long* dbg_buf = nullptr;
    static int dbg_on = -1;
    if (dbg_on < 0) { const char* e = getenv("MI_B200_DEBUG_TIMING"); dbg_on = (e && e[0] == '1') ? 1 : 0; }
    kp.dbg = nullptr;
    if (dbg_on) {
        if (!dbg_buf) cudaMalloc(&dbg_buf, 16 * sizeof(unsigned long long));
        cudaMemsetAsync(dbg_buf, 0, 16 * sizeof(unsigned long long), stream);
        kp.dbg = dbg_buf;
    }
    mi_prof_begin(MI_TAG_FPROP_KXS, mi_conv_flops(n, h, wd, cin, cout, 3),
                  mi_conv_bytes(n, h, wd, cin, cout, 3), stream);
    kernels[stream_w ? 1 : 0][epi]<<<grid, KXS_THREADS, smem, stream>>>(map_x, map_w, map_y, map_op, kp);
    mi_prof_end(stream);
    if (dbg_on) {
        unsigned long long d[16];
        cudaStreamSynchronize(stream);
        cudaMemcpy(d, dbg_buf, sizeof(d), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[kxs%s cta0] items=%llu producer: wait_empty=%llu total=%llu | mma: wait_weights=%llu "
                "wait_full=%llu wait_tmem_empty=%llu total=%llu | epilogue: wait_acc=%llu tmem_ld=%llu "
                "math+store=%llu total=%llu cycles (stages=%d bn=%d chunks=%d n_tiles=%d grid=%d epi=%d) | mma issue loops: "
                "%llu cycles over %llu stages | epilogue leader: store wait_read=%llu staging wait=%llu\n",
                stream_w ? "-stream" : "", d[10], d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7], d[8], d[9], kp.stages,
                kp.bn, kp.chunks, kp.n_tiles, grid, epi, d[11], d[12], d[13], d[14]);
    }
    MI_LAUNCHED();
    MI_RETURN_LAST();
}

}  // namespace

bool mi_tc_fprop_eligible(const float* x, int ldx, const float* w, int ldw, const float* y, int ldy, int n, int h,
                          int wd, int cin, int cout, int k) {
    (void)y; (void)ldy; (void)n;
    if (!device_is_sm100() || !encode_fn()) return false;
    if (!aligned_view(x, ldx) || !aligned_view(w, ldw)) return false;
    if (k > 7 || wd < 8 || h < 1) return false;
    if (cout < 16) return false;   // tiny heads stay on the exact fp32 SIMT engine
    if (cin < 4) return false;     // (a 32-channel TMA box over a 6-channel tensor is legal: the tail is zero-filled)
    return true;
}

int mi_tc_fprop(const float* x, int ldx, const float* w, int ldw, const float* bias, float* y, int ldy,
                const float* mask_y, int ldmask, int mask_act, float mask_slope, int accumulate, int n, int h, int wd,
                int cin, int cout, int k, int act, float slope, cudaStream_t stream) {
    if (k == 3 && cout <= 512 && kxs_enabled(n, h, wd)) {
        const int rc = launch_kxs(x, ldx, w, ldw, bias, y, ldy, mask_y, ldmask, mask_act, mask_slope, accumulate, n, h, wd,
                                  cin, cout, act, slope, stream);
        if (rc != MI_ERR_UNSUPPORTED) return rc;
    }
    if (k == 3 && cin <= 64 && cout <= 64 && halo_enabled()) {
        HaloParams hp;
        hp.n = n; hp.h = h; hp.w = wd; hp.cin = cin; hp.cout = cout;
        hp.chunks = mi_cdiv(cin, KCH);
        hp.bn = cout <= 32 ? 32 : 64;
        hp.tiles_x = mi_cdiv(wd, HT_W);
        hp.tiles_y = mi_cdiv(h, HT_H);
        hp.total_tiles = hp.tiles_x * hp.tiles_y * n;
        hp.act = act; hp.slope = slope; hp.accumulate = accumulate; hp.mask_act = mask_act; hp.rnd = mi_tf32_rn_enabled();
        hp.mask_slope = mask_slope; hp.ldy = ldy; hp.ldmask = ldmask; hp.bias = bias; hp.mask_y = mask_y; hp.y = y;
        const size_t b_total = (size_t)9 * hp.chunks * hp.bn * ROW_BYTES;
        const size_t budget = 227 * 1024 - 22 * 1024 - 2048;   // 227 KB per CTA minus static smem (epilogue tiles, bias)
        hp.halo_w = halo_pitch();
        hp.halo_bytes = (uint32_t)hp.halo_w * HALO_H * ROW_BYTES;
        hp.halo_stride = (hp.halo_bytes + 1023u) & ~1023u;
        int stages = (int)((budget - b_total) / hp.halo_stride);
        if (stages > 6) stages = 6;
        if (stages >= 2) {
            hp.stages = stages;
            const size_t smem = b_total + (size_t)stages * hp.halo_stride + (2 * stages + 7) * 8 + 1024;
            CUtensorMap map_x, map_w;
            if (!make_act_map(&map_x, x, ldx, n, h, wd, cin, hp.halo_w, HALO_H)) return MI_ERR_UNSUPPORTED;
            if (!make_weight_map(&map_w, w, ldw, cout, 9, cin, hp.bn)) return MI_ERR_UNSUPPORTED;
            static bool attr = false;
            if (!attr) {
                cudaError_t e = cudaFuncSetAttribute(conv_fprop_tc_halo_kernel,
                                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(206 * 1024));
                if (e != cudaSuccess) return (int)e;
                attr = true;
            }
            int grid = hp.total_tiles < num_sms() ? hp.total_tiles : num_sms();
            mi_prof_begin(MI_TAG_FPROP_HALO, mi_conv_flops(n, h, wd, cin, cout, k), mi_conv_bytes(n, h, wd, cin, cout, k),
                          stream);
            static unsigned long long* dbg_buf = nullptr;
            static int dbg_on = -1;
            if (dbg_on < 0) { const char* e = getenv("MI_B200_DEBUG_TIMING"); dbg_on = (e && e[0] == '1') ? 1 : 0; }
            hp.dbg = nullptr;
            if (dbg_on) {
                if (!dbg_buf) cudaMalloc(&dbg_buf, 16 * sizeof(unsigned long long));
                cudaMemsetAsync(dbg_buf, 0, 16 * sizeof(unsigned long long), stream);
                hp.dbg = dbg_buf;
            }
            conv_fprop_tc_halo_kernel<<<grid, HALO_THREADS, smem, stream>>>(map_x, map_w, hp);
            if (dbg_on) {
                unsigned long long h[16];
                cudaStreamSynchronize(stream);
                cudaMemcpy(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost);
                fprintf(stderr, "[halo cta0] tiles=%llu producer: wait_empty=%llu total=%llu | mma: wait_weights=%llu "
                        "wait_full=%llu wait_tmem_empty=%llu total=%llu | epilogue: wait_acc=%llu tmem_ld=%llu "
                        "math+store=%llu total=%llu cycles (stages=%d bn=%d chunks=%d)\n", h[10], h[0], h[1], h[2], h[3],
                        h[4], h[5], h[6], h[7], h[8], h[9], hp.stages, hp.bn, hp.chunks);
            }
            mi_prof_end(stream);
            MI_LAUNCHED();
            MI_RETURN_LAST();
        }
    }
    if (k == 3 && (cin > 64 || cout > 64) && cout <= 512 && halo_stream_enabled()) {
        HaloStreamParams hp;
        hp.n = n; hp.h = h; hp.w = wd; hp.cin = cin; hp.cout = cout;
        hp.chunks = mi_cdiv(cin, KCH);
        hp.tiles_x = mi_cdiv(wd, HT_W);
        hp.tiles_y = mi_cdiv(h, HT_H);
        hp.total_tiles = hp.tiles_x * hp.tiles_y * n;
        hp.n_tiles = mi_cdiv(cout, HS_BN);
        hp.items = hp.total_tiles * hp.n_tiles;
        hp.halo_w = halo_pitch();
        hp.halo_bytes = (uint32_t)hp.halo_w * HALO_H * ROW_BYTES;
        hp.halo_stride = (hp.halo_bytes + 1023u) & ~1023u;
        hp.act = act; hp.slope = slope; hp.accumulate = accumulate; hp.mask_act = mask_act; hp.rnd = mi_tf32_rn_enabled();
        hp.mask_slope = mask_slope; hp.ldy = ldy; hp.ldmask = ldmask; hp.bias = bias; hp.mask_y = mask_y; hp.y = y;
        const size_t stage_bytes = (size_t)hp.halo_stride + 9 * (size_t)HS_BN * ROW_BYTES;
        const size_t smem = HS_STAGES * stage_bytes + (2 * HS_STAGES + 5) * 8 + 1024;
        if (smem <= 205 * 1024) {
            CUtensorMap map_x, map_w;
            if (!make_act_map(&map_x, x, ldx, n, h, wd, cin, hp.halo_w, HALO_H)) return MI_ERR_UNSUPPORTED;
            if (!make_weight_map(&map_w, w, ldw, cout, 9, cin, HS_BN)) return MI_ERR_UNSUPPORTED;
            static bool attr = false;
            if (!attr) {
                cudaError_t e = cudaFuncSetAttribute(conv_fprop_tc_halo_stream_kernel,
                                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(205 * 1024));
                if (e != cudaSuccess) return (int)e;
                attr = true;
            }
            const int grid = hp.items < num_sms() ? hp.items : num_sms();
            static unsigned long long* dbg_buf = nullptr;
            static int dbg_on = -1;
            if (dbg_on < 0) { const char* e = getenv("MI_B200_DEBUG_TIMING"); dbg_on = (e && e[0] == '1') ? 1 : 0; }
            hp.dbg = nullptr;
            if (dbg_on) {
                if (!dbg_buf) cudaMalloc(&dbg_buf, 16 * sizeof(unsigned long long));
                cudaMemsetAsync(dbg_buf, 0, 16 * sizeof(unsigned long long), stream);
                hp.dbg = dbg_buf;
            }
            mi_prof_begin(MI_TAG_FPROP_STREAM, mi_conv_flops(n, h, wd, cin, cout, k), mi_conv_bytes(n, h, wd, cin, cout, k),
                          stream);
            conv_fprop_tc_halo_stream_kernel<<<grid, HALO_THREADS, smem, stream>>>(map_x, map_w, hp);
            mi_prof_end(stream);
            if (dbg_on) {
                unsigned long long d[16];
                cudaStreamSynchronize(stream);
                cudaMemcpy(d, dbg_buf, sizeof(d), cudaMemcpyDeviceToHost);
                fprintf(stderr, "[stream cta0] items=%llu n_tiles=%d chunks=%d grid=%d | producer: wait_empty=%llu total=%llu | "
                        "mma: wait_full=%llu wait_tmem_empty=%llu first_data_at=%llu total=%llu | epilogue: wait_acc=%llu "
                        "total=%llu cycles\n", d[10], hp.n_tiles, hp.chunks, grid, d[0], d[2], d[3], d[5], d[6], d[7], d[8],
                        d[9]);
            }
            MI_LAUNCHED();
            MI_RETURN_LAST();
        }
    }
    if ((k == 5 || k == 7) && cout <= 512 && halo_stream_enabled()) {
        HaloRowsParams hp;
        hp.n = n; hp.h = h; hp.w = wd; hp.cin = cin; hp.cout = cout; hp.k = k;
        hp.chunks = mi_cdiv(cin, KCH);
        hp.bn = cout <= 32 ? 32 : 64;
        hp.tiles_x = mi_cdiv(wd, HT_W);
        hp.tiles_y = mi_cdiv(h, HT_H);
        hp.total_tiles = hp.tiles_x * hp.tiles_y * n;
        hp.n_tiles = mi_cdiv(cout, hp.bn);
        hp.items = hp.total_tiles * hp.n_tiles;
        hp.halo_w = HT_W + k - 1;
        hp.halo_bytes = (uint32_t)hp.halo_w * (HT_H + k - 1) * ROW_BYTES;
        hp.halo_stride = (hp.halo_bytes + 1023u) & ~1023u;
        hp.sa = 2;
        const size_t b_stage = (size_t)k * hp.bn * ROW_BYTES;
        int sb = (int)((200 * 1024 - (size_t)hp.sa * hp.halo_stride) / b_stage);
        if (sb > 4) sb = 4;
        hp.sb = sb;
        hp.act = act; hp.slope = slope; hp.accumulate = accumulate; hp.mask_act = mask_act; hp.rnd = mi_tf32_rn_enabled();
        hp.mask_slope = mask_slope; hp.ldy = ldy; hp.ldmask = ldmask; hp.bias = bias; hp.mask_y = mask_y; hp.y = y;
        if (sb >= 2) {
            const size_t smem = (size_t)hp.sa * hp.halo_stride + (size_t)sb * b_stage + (2 * hp.sa + 2 * sb + 5) * 8 + 1024;
            CUtensorMap map_x, map_w;
            if (!make_act_map(&map_x, x, ldx, n, h, wd, cin, hp.halo_w, HT_H + k - 1)) return MI_ERR_UNSUPPORTED;
            if (!make_weight_map(&map_w, w, ldw, cout, k * k, cin, hp.bn)) return MI_ERR_UNSUPPORTED;
            static bool attr = false;
            if (!attr) {
                cudaError_t e = cudaFuncSetAttribute(conv_fprop_tc_halo_rows_kernel,
                                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(205 * 1024));
                if (e != cudaSuccess) return (int)e;
                attr = true;
            }
            const int grid = hp.items < num_sms() ? hp.items : num_sms();
            mi_prof_begin(MI_TAG_FPROP_STREAM, mi_conv_flops(n, h, wd, cin, cout, k), mi_conv_bytes(n, h, wd, cin, cout, k),
                          stream);
            conv_fprop_tc_halo_rows_kernel<<<grid, HALO_THREADS, smem, stream>>>(map_x, map_w, hp);
            mi_prof_end(stream);
            MI_LAUNCHED();
            MI_RETURN_LAST();
        }
    }
    FpropParams p;
    p.n = n; p.h = h; p.w = wd; p.cin = cin; p.cout = cout; p.k = k;
    pick_tile(wd, BM, &p.tw, &p.th);
    p.tiles_x = mi_cdiv(wd, p.tw);
    p.tiles_y = mi_cdiv(h, p.th);
    p.bn = pick_bn(cout);
    // Each SM fills its shared memory from L2 at a fixed rate (~130 GB/s measured), so a deep layer with few
    // 128-pixel tiles is bound by how many SMs take part: shrink the channel tile until the grid covers the chip.
    while (p.bn > 64 && (long long)p.tiles_x * p.tiles_y * n * mi_cdiv(cout, p.bn) < num_sms()) p.bn >>= 1;
    p.act = act; p.slope = slope; p.accumulate = accumulate; p.mask_act = mask_act; p.mask_slope = mask_slope;
    p.rnd = mi_tf32_rn_enabled();
    p.ldy = ldy; p.ldmask = ldmask; p.bias = bias; p.mask_y = mask_y; p.y = y;
    const size_t stage_bytes = (size_t)(BM + p.bn) * ROW_BYTES;
    int stages = (int)((200 * 1024) / stage_bytes);
    if (stages > 6) stages = 6;
    // leave room for 2 CTAs per SM so epilogues overlap mainloops -- unless the grid cannot put two CTAs on an SM
    // anyway (deep layers: 32-96 CTAs with 144-iteration K loops): there the pipeline depth is what hides the
    // L2 latency (3 stages x 24 KB in flight measured ~50 GB/s per SM), so those take every stage that fits
    const long long ctas = (long long)p.tiles_x * p.tiles_y * n * mi_cdiv(cout, p.bn);
    if (p.bn <= 128 && stages > 3 && ctas > num_sms()) stages = 3;
    if (stages < 2) stages = 2;
    p.stages = stages;
    const size_t smem = stages * stage_bytes + (2 * stages + 2) * 8 + 1024;
    CUtensorMap map_x, map_w;
    if (!make_act_map(&map_x, x, ldx, n, h, wd, cin, p.tw, p.th)) return MI_ERR_UNSUPPORTED;
    if (!make_weight_map(&map_w, w, ldw, cout, k * k, cin, p.bn)) return MI_ERR_UNSUPPORTED;
    static size_t smem_set = 0;
    if (smem > smem_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_fprop_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)(220 * 1024));
        if (e != cudaSuccess) return (int)e;
        smem_set = 220 * 1024;
    }
    dim3 grid(p.tiles_x * p.tiles_y * n, mi_cdiv(cout, p.bn));
    mi_prof_begin(MI_TAG_FPROP_TC, mi_conv_flops(n, h, wd, cin, cout, k), mi_conv_bytes(n, h, wd, cin, cout, k), stream);
    conv_fprop_tc_kernel<<<grid, NTHREADS, smem, stream>>>(map_x, map_w, p);
    mi_prof_end(stream);
    MI_LAUNCHED();
    MI_RETURN_LAST();
}

// every 3x3 / 5x5 / 7x7 layer takes the filter-column kernel (MI_B200_WGRAD_KX=0 keeps the per-tap kernel: A/B switch)
bool mi_tc_wgrad_kx_shape(int cin, int cout, int k) {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("MI_B200_WGRAD_KX");
        on = (e && e[0] == '0') ? 0 : 1;
    }
    (void)cout;
    // 5x5 / 7x7: k filter rows stacked along N (k * 32 columns per 32-channel X box); the two boxes of a 64-cin block
    // need 2 * k * 32 <= 512 TMEM columns
    return on && (k == 3 || k == 5 || k == 7) && (long long)(cin >= 64 ? 2 : 1) * k * 32 <= 512 && device_is_sm100() &&
           encode_fn();
}

// Cout <= 64: two filter columns share a CTA (the halves of the M = 128 accumulator); MI_B200_WGRAD_PAIR=0: A/B switch
bool mi_tc_wgrad_kx_pair(int cin, int cout, int k) {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("MI_B200_WGRAD_PAIR");
        on = (e && e[0] == '0') ? 0 : 1;
    }
    return on && cout <= 64 && mi_tc_wgrad_kx_shape(cin, cout, k);
}

bool mi_tc_wgrad_eligible(const float* x, int ldx, const float* dy, int lddy, int n, int h, int wd, int cin, int cout,
                          int k) {
    (void)n;
    if (!device_is_sm100() || !encode_fn()) return false;
    if (!aligned_view(x, ldx) || !aligned_view(dy, lddy)) return false;
    if (k > 7 || wd < 8 || h < 1) return false;
    if (cout < 16 || cin < 4) return false;
    return true;
}

int mi_tc_wgrad_partials(const float* x, int ldx, const float* dy, int lddy, int n, int h, int wd, int cin, int cout,
                         int k, int ldw, float* ws_w, float* ws_b, int splits, int* bias_splits_out,
                         cudaStream_t stream) {
    *bias_splits_out = mi_bias_splits((long long)n * h * wd);
    if (mi_tc_wgrad_kx_shape(cin, cout, k)) {
        WgradKxParams q;
        q.n = n; q.h = h; q.w = wd; q.cin = cin; q.cout = cout; q.k = k; q.ldw = ldw;
        q.na = cout >= BM ? 4 : mi_cdiv(cout, KCH);      // dY boxes per CTA (a partial last cout tile is zero-filled)
        q.nb = cin >= 64 ? 2 : mi_cdiv(cin, KCH);        // X boxes per CTA
        q.ci_tiles = mi_cdiv(cin, 64);
        const int co_tiles = mi_cdiv(cout, BM);
        q.ws_w = ws_w;
        static int bias_fused = -1;      // MI_B200_WGRAD_BIAS_FUSED=0: separate bias_partial_kernel pass (A/B switch)
        if (bias_fused < 0) { const char* e = getenv("MI_B200_WGRAD_BIAS_FUSED"); bias_fused = (e && e[0] == '0') ? 0 : 1; }
        q.ws_b = bias_fused ? ws_b : nullptr;
        q.pair = mi_tc_wgrad_kx_pair(cin, cout, k) ? 1 : 0;
        const int a_slots = q.pair ? 2 * q.na : q.na;
        const size_t row_bytes = 8 * ROW_BYTES;
        // Tile height: every pipeline stage costs the MMA lane one barrier wait (~450 cycles), so the narrow layers
        // (one or two boxes per operand: 8-16 MMAs per 8-row stage) take 16-row tiles when three stages still fit.
        // (the split count was sized on 8x8 tiles; with 16 rows a CTA simply walks half as many, twice as large)
        int rows = 16, stages = 0;
        size_t stage_bytes = 0, tail = 0;
        static int min16 = -1;           // MI_B200_WGRAD_MIN16: stages a 16-row tile must leave room for (default 3)
        if (min16 < 0) { const char* e = getenv("MI_B200_WGRAD_MIN16"); min16 = e ? atoi(e) : 3; }
        const int min_stages16 = min16;
        for (;; rows = 8) {
            stage_bytes = (size_t)a_slots * rows * row_bytes + (size_t)q.nb * (rows + k - 1) * row_bytes;
            tail = (size_t)(4 - a_slots) * rows * row_bytes;
            stages = (int)((200 * 1024 - tail) / stage_bytes);
            if (stages > 6) stages = 6;
            if (rows == 8 || (stages >= min_stages16 && h >= 16)) break;
        }
        if (stages < 2) return MI_ERR_UNSUPPORTED;
        q.rows = rows;
        q.stages = stages;
        q.tiles_x = mi_cdiv(wd, 8);
        q.tiles_y = mi_cdiv(h, q.rows);
        q.total_tiles = q.tiles_x * q.tiles_y * n;
        q.tiles_per_split = mi_cdiv(q.total_tiles, splits);
        const size_t smem = stages * stage_bytes + tail + (2 * stages + 2) * 8 + 1024;
        CUtensorMap map_dy, map_x;
        if (!make_act_map(&map_dy, dy, lddy, n, h, wd, cout, 8, q.rows, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))
            return MI_ERR_UNSUPPORTED;
        static int merge_on = -1;        // MI_B200_WGRAD_MERGED=0: two N = 96 MMAs per tile row everywhere (A/B switch)
        if (merge_on < 0) { const char* e = getenv("MI_B200_WGRAD_MERGED"); merge_on = (e && e[0] == '0') ? 0 : 1; }
        q.merged = (merge_on && k == 3 && q.nb == 2 && cin % 64 == 0) ? 1 : 0;
        if (q.merged) {
            // x[n][y][x][half][c32]: the channel dimension split in two so that one box brings both halves of a tile row
            EncodeTiledFn fn = encode_fn();
            cuuint64_t dims[5] = {(cuuint64_t)KCH, (cuuint64_t)wd, (cuuint64_t)(cin / KCH), (cuuint64_t)h, (cuuint64_t)n};
            cuuint64_t strides[4] = {(cuuint64_t)ldx * 4, (cuuint64_t)KCH * 4, (cuuint64_t)wd * ldx * 4,
                                     (cuuint64_t)h * wd * ldx * 4};
            cuuint32_t box[5] = {(cuuint32_t)KCH, 8, 2, (cuuint32_t)(q.rows + k - 1), 1};
            cuuint32_t estr[5] = {1, 1, 1, 1, 1};
            if (!fn || fn(&map_x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(x), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                q.merged = 0;            // the driver refused the 5-D view: two 4-D boxes as before
        }
        if (!q.merged &&
            !make_act_map(&map_x, x, ldx, n, h, wd, cin, 8, q.rows + k - 1, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))
            return MI_ERR_UNSUPPORTED;
        static bool attr_kx = false;
        if (!attr_kx) {
            cudaError_t e = cudaFuncSetAttribute(conv_wgrad_tc_kx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)(220 * 1024));
            if (e != cudaSuccess) return (int)e;
            attr_kx = true;
        }
        dim3 grid((q.pair ? (k + 1) / 2 : k) * q.ci_tiles * co_tiles, splits);
        mi_prof_begin(MI_TAG_WGRAD_KX, mi_conv_flops(n, h, wd, cin, cout, k), mi_conv_bytes(n, h, wd, cin, cout, k), stream);
        conv_wgrad_tc_kx_kernel<<<grid, NTHREADS, smem, stream>>>(map_dy, map_x, q);
        mi_prof_end(stream);
        MI_LAUNCHED();
        cudaError_t e = cudaPeekAtLastError();
        if (e != cudaSuccess) return (int)e;
        if (q.ws_b) {                    // one bias partial per split-K slice came out of the kernel itself
            *bias_splits_out = splits;
            return MI_OK;
        }
        const long long m_total = (long long)n * h * wd;
        const int bsplits = mi_bias_splits(m_total);
        const long long chunk = (m_total + bsplits - 1) / bsplits;
        const int vec = mi_al16(dy) && (lddy % 4 == 0) && ((cout % 4 == 0) || lddy == ((cout + 3) & ~3));
        bias_partial_kernel<<<dim3(bsplits, mi_cdiv(cout, 32)), 256, 0, stream>>>(dy, lddy, ws_b, cout, m_total, chunk, vec);
        MI_LAUNCHED();
        *bias_splits_out = bsplits;
        MI_RETURN_LAST();
    }
    WgradParams p;
    p.n = n; p.h = h; p.w = wd; p.cin = cin; p.cout = cout; p.k = k; p.ldw = ldw;
    pick_tile(wd, 64, &p.pw, &p.ph);
    p.tiles_x = mi_cdiv(wd, p.pw);
    p.tiles_y = mi_cdiv(h, p.ph);
    p.total_tiles = p.tiles_x * p.tiles_y * n;
    p.bn = cin <= 32 ? 32 : (cin <= 64 ? 64 : 128);
    p.co_tiles = mi_cdiv(cout, BM);
    p.ci_tiles = mi_cdiv(cin, p.bn);
    p.tiles_per_split = mi_cdiv(p.total_tiles, splits);
    p.ws_w = ws_w;
    const size_t stage_bytes = (size_t)(BM / KCH + p.bn / KCH) * 64 * ROW_BYTES;
    int stages = (int)((200 * 1024) / stage_bytes);
    if (stages > 4) stages = 4;
    if (stages < 2) stages = 2;
    p.stages = stages;
    const size_t smem = stages * stage_bytes + (2 * stages + 2) * 8 + 1024;
    CUtensorMap map_dy, map_x;
    if (!make_act_map(&map_dy, dy, lddy, n, h, wd, cout, p.pw, p.ph, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))
        return MI_ERR_UNSUPPORTED;
    if (!make_act_map(&map_x, x, ldx, n, h, wd, cin, p.pw, p.ph, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))
        return MI_ERR_UNSUPPORTED;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(conv_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)(220 * 1024));
        if (e != cudaSuccess) return (int)e;
        attr = true;
    }
    dim3 grid(k * k * p.co_tiles * p.ci_tiles, splits);
    mi_prof_begin(MI_TAG_WGRAD_TC, mi_conv_flops(n, h, wd, cin, cout, k), mi_conv_bytes(n, h, wd, cin, cout, k), stream);
    conv_wgrad_tc_kernel<<<grid, NTHREADS, smem, stream>>>(map_dy, map_x, p);
    mi_prof_end(stream);
    MI_LAUNCHED();
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) return (int)e;
    const long long m_total = (long long)n * h * wd;
    const int bsplits = mi_bias_splits(m_total);
    const long long chunk = (m_total + bsplits - 1) / bsplits;
    // float4 lanes may read the pad lane of a 4-padded row (never a neighbouring concat slice)
    const int vec = mi_al16(dy) && (lddy % 4 == 0) && ((cout % 4 == 0) || lddy == ((cout + 3) & ~3));
    bias_partial_kernel<<<dim3(bsplits, mi_cdiv(cout, 32)), 256, 0, stream>>>(dy, lddy, ws_b, cout, m_total, chunk, vec);
    MI_LAUNCHED();
    MI_RETURN_LAST();
}

extern "C" int mi_tc_available(void) { return (device_is_sm100() && encode_fn()) ? 1 : 0; }

extern "C" int mi_set_pad_lanes_scratch(int on) {
    const int prev = g_pad_lanes_scratch;
    g_pad_lanes_scratch = on ? 1 : 0;
    return prev;
}
