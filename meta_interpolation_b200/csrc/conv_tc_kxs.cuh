// 3x3 fprop / dgrad with the three filter COLUMNS stacked along the MMA N dimension (included by conv_tc.cu inside its
// anonymous namespace, after the PTX wrappers and the epilogue helpers).
//
// Why: in SS mode a `tcgen05.mma kind::tf32` of 128 x N x 8 does not get cheaper below N ~ 128 on B200 -- the
// 128 x 32-byte A-operand fetch from shared memory sets a floor (per-role cycle counters: 60-72 cycles per MMA at
// N = 64 in the halo kernels, ~100 at N = 192 here, against a tensor-pipe floor of N / 2).  The halo kernels issue nine
// N = Cout <= 64 MMAs per 8-channel K slice, i.e. they run the tensor pipe at <= 50 % by construction (ncu: 32.5 %).
// Here ONE MMA serves three taps (ncu: 47-49 % on a 148-CTA launch, 59-61 % at the 37-CTA share of the 8-lane run):
//
//     D[p][(kx, co)] += sum_ci  Xbox[row(p) + ky][col(p)][ci] * W[ky][kx][ci][co]          (N = 3 * Cout-tile = 192)
//
// over an 8-row x 16-column pixel tile p (M = 128) of a 10 x 16 input box, so a K slice needs three MMAs (ky) instead of
// nine and every A fetch is shared by three taps.  The accumulator then holds, per filter column kx, the partial output
// that belongs to the pixel kx - 1 columns to the RIGHT of where it sits, and the epilogue finishes the convolution
// with a one-lane shift:
//
//     out[r][j] = D[r][j-1][kx=0] + D[r][j][kx=1] + D[r][j+1][kx=2],      j = 1..14  (columns 0 and 15 are halo)
//
// A tile row is 16 consecutive TMEM lanes and a warp reads two whole rows, so the shift is a __shfl_up / __shfl_down
// by one lane that never crosses a row or a warp.  14 of 16 columns produce output (87.5 %); rows need no halo in the
// M dimension because ky is an operand VIEW (the box is 16 pixels = two 1024-byte swizzle atoms wide, so the view of
// filter row ky starts ky * 2048 bytes into the box, atom-aligned, and all 128 pixel rows are contiguous: SBO = 1024).
//
// STREAM = false: all weights of the layer (Cin, Cout <= 64) stay resident in shared memory.
// STREAM = true : > 64 channels; a pipeline stage carries the box of one 32-channel chunk and the nine
//                 {64 cout x 32 cin} weight tiles of that chunk; persistent CTAs walk (pixel tile, 64-cout tile) items.
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer, warps 2-9 epilogue (two per TMEM lane quarter, 16 columns each),
// warp 10 output stores (see KXS_THREADS).  The epilogue releases the accumulator buffer as soon as its TMEM reads
// are done, before the activation / staging / stores.

constexpr int KX_W = 16, KX_OW = 14, KX_H = 8, KX_BOX_H = 10;
constexpr uint32_t KX_BOX_BYTES = KX_W * KX_BOX_H * ROW_BYTES;       // 20 KB, a multiple of 1024
constexpr uint32_t KX_ROW_VIEW = KX_W * ROW_BYTES;                   // distance between the views of ky and ky + 1
constexpr uint32_t KX_OUT_BOX = KX_OW * KX_H * ROW_BYTES;            // one staged {32 ch, 14, 8} output box: 14 KB
constexpr uint32_t KX_STAGING = 2 * KX_OUT_BOX;                      // two 32-channel boxes per 64-cout tile

// Epilogue modes (template parameter EPI).  ncu on the first version of this kernel -- the per-thread staged-transpose
// stores of the halo kernels, ~10k SASS instructions -- showed the eight epilogue warps stalled on instruction fetch
// (stall_no_inst the top reason) and ~4000 cycles per tile against ~3000 for the MMAs.  The specialised modes keep the
// per-tile code to a few dozen instructions per thread:
//   0  plain            registers -> 128B-swizzled staging tile -> ONE cp.async.bulk.tensor store per 32-channel box
//                       (TMA clips image edges and the channel tail, so no per-thread bounds or address arithmetic)
//   1  operand in place the activation-derivative mask (dgrad into an activated tensor) or the previous value of y
//                       (accumulate) is TMA-LOADED into the staging tile while the MMAs of the tile still run; each
//                       thread combines its own 64 bytes in place; then the same bulk store
//   2  generic          unaligned views, or mask and accumulate together: the per-thread path of the halo kernels
enum { KXS_EPI_PLAIN = 0, KXS_EPI_OPERAND = 1, KXS_EPI_GENERIC = 2 };
// warp 0 producer, warp 1 MMA, warps 2-9 epilogue, warp 10 output stores.  The store warp exists because the thread that
// issues a bulk store has to wait until TMA has READ the staging tile before anyone may overwrite it (~1400 cycles
// for two boxes, per-role counters); while an epilogue warp did that, its share of the NEXT accumulator stayed unread
// in TMEM and the MMA lane waited for the buffer.
constexpr int KXS_THREADS = 352;

struct KxsParams {
    int rnd;
    int n, h, w, cin, cout, chunks, bn, stages, n_tiles, tiles_x, tiles_y, total_tiles, items, act, accumulate, mask_act,
        ldy, ldmask, c_tma,             // c_tma: channels the TMA store / operand load covers (see the epilogue)
        one_box;                        // plain epilogue through ONE 14 KB staging box, freeing room for a third stage
    unsigned long long* dbg;            // optional per-role cycle counters of CTA 0 (MI_B200_DEBUG_TIMING=1)
    float slope, mask_slope;
    const float* bias;
    const float* mask_y;
    float* y;
};

__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read_but_one() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// named barriers between the 8 epilogue warps and the store warp (288 threads): one side arrives, the other waits
__device__ __forceinline__ void kxs_bar_sync(int id) { asm volatile("bar.sync %0, 288;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void kxs_bar_arrive(int id) { asm volatile("bar.arrive %0, 288;" ::"r"(id) : "memory"); }

// sigmoid / tanh epilogues (a handful of head layers): out of line, so the hot epilogue stays small
__device__ __noinline__ float kxs_act_rare(float v, int act, float slope) { return mi_act_apply(v, act, slope); }
__device__ __noinline__ float kxs_act_grad_rare(float y, int act, float slope) { return mi_act_grad(y, act, slope); }

// three MMAs (filter rows) per 8-channel K slice; KS = K slices of this chunk that hold data
#define MI_KXS_MMAS(KS)                                                                                             \
    _Pragma("unroll") for (int ky = 0; ky < 3; ++ky) {                                                              \
        const uint64_t a_ky = desc_advance(ad0, (uint32_t)ky * KX_ROW_VIEW);                                        \
        const uint64_t b_ky = desc_advance(bd0, (uint32_t)ky * 3u * b_tile);                                        \
        _Pragma("unroll") for (int kk = 0; kk < KS; ++kk)                                                           \
            umma_tf32(d_addr, desc_advance(a_ky, kk * 32), desc_advance(b_ky, kk * 32), idesc,                      \
                      (ch > 0 || ky > 0 || kk > 0) ? 1u : 0u);                                                      \
    }

template <bool STREAM, int EPI>
__global__ void __launch_bounds__(KXS_THREADS)
conv_fprop_tc_kxs_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                         const __grid_constant__ CUtensorMap map_y, const __grid_constant__ CUtensorMap map_op,
                         const KxsParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t b_tile = (uint32_t)p.bn * ROW_BYTES;            // one (tap, chunk) weight tile
    const uint32_t b_chunk = 9u * b_tile;                          // the nine taps of a 32-channel chunk, tap-major
    // resident: [weights: chunks * b_chunk][S boxes]; streamed: S x [box][b_chunk]
    const uint32_t stage_bytes = STREAM ? KX_BOX_BYTES + b_chunk : KX_BOX_BYTES;
    uint8_t* smem_a = STREAM ? smem : smem + (size_t)p.chunks * b_chunk;
    const int S = p.stages;
    uint8_t* staging = smem_a + (size_t)S * stage_bytes;            // epilogue tile(s), 1024-aligned
    uint64_t* bars = reinterpret_cast<uint64_t*>(staging + (p.one_box ? KX_OUT_BOX : KX_STAGING));
    // bars: [0,S) full, [S,2S) empty, 2S..2S+1 tmem full[2], 2S+2..2S+3 tmem empty[2], 2S+4.. weights of chunk c (resident),
    //       2S+6 = epilogue operand (mask / previous y) landed in the staging tile
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 4 + 3);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nstack = 3 * p.bn;                                   // MMA N and accumulator columns per buffer
    const uint32_t tmem_cols = 2 * nstack <= 256 ? 256u : 512u;

    __shared__ float sbias[512];
    for (int i = threadIdx.x; i < 512; i += KXS_THREADS) sbias[i] = (p.bias && i < p.cout) ? p.bias[i] : 0.f;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(smem_u32(&bars[s]), 1);
            mbar_init(smem_u32(&bars[S + s]), 1);
        }
        mbar_init(smem_u32(&bars[2 * S]), 1);
        mbar_init(smem_u32(&bars[2 * S + 1]), 1);
        mbar_init(smem_u32(&bars[2 * S + 2]), 256);     // all eight epilogue warps release an accumulator buffer
        mbar_init(smem_u32(&bars[2 * S + 3]), 256);
        mbar_init(smem_u32(&bars[2 * S + 4]), 1);
        mbar_init(smem_u32(&bars[2 * S + 5]), 1);
        mbar_init(smem_u32(&bars[2 * S + 6]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int my_items = (p.items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (warp == 0) {
        // resident weights: one barrier per 32-channel chunk, the first box in between (the MMAs of chunk 0 start
        // before the second half of the filter bank has landed)
        auto load_weights = [&](int ch) {
            const uint32_t wbar = smem_u32(&bars[2 * S + 4 + ch]);
            mbar_expect_tx(wbar, b_chunk);
            for (int tap = 0; tap < 9; ++tap)
                tma_load_3d(smem_u32(smem) + (uint32_t)ch * b_chunk + (uint32_t)tap * b_tile, &map_w, wbar, ch * KCH, tap, 0);
        };
        if (!STREAM) {
            if (elect_one()) load_weights(0);
            __syncwarp();
        }
        int it = 0;
        long long t_empty = 0;
        const long long t_begin = clock64();
        for (int t = 0; t < my_items; ++t) {
            const int item = (int)blockIdx.x + t * (int)gridDim.x;
            const int nt = STREAM ? item % p.n_tiles : 0;
            int tile = STREAM ? item / p.n_tiles : item;
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
                    const uint32_t base = smem_u32(smem_a + (size_t)s * stage_bytes);
                    mbar_expect_tx(full, STREAM ? KX_BOX_BYTES + b_chunk : KX_BOX_BYTES);
                    tma_load_4d(base, &map_x, full, ch * KCH, tx_i * KX_OW - 1, ty_i * KX_H - 1, img);
                    if (STREAM) {
#pragma unroll
                        for (int tap = 0; tap < 9; ++tap)
                            tma_load_3d(base + KX_BOX_BYTES + (uint32_t)tap * b_tile, &map_w, full, ch * KCH, tap,
                                        nt * p.bn);
                    } else if (it == 0 && p.chunks > 1) {
                        load_weights(1);
                    }
                }
                __syncwarp();
            }
        }
        if (!STREAM && my_items == 0 && p.chunks > 1) {
            // (a CTA without tiles still owes the second barrier its bytes before the block may retire)
            if (elect_one()) load_weights(1);
            __syncwarp();
        }
        if (p.dbg && blockIdx.x == 0 && lane == 0) {
            p.dbg[0] = (unsigned long long)t_empty; p.dbg[1] = (unsigned long long)(clock64() - t_begin);
        }
    } else if (warp == 1) {
        // The issuing lane's work BETWEEN two stages is on the critical path: tcgen05.mma blocks at issue once a couple
        // of instructions are queued, so the tensor pipe runs dry while this warp walks from the last MMA of a stage
        // to the first of the next (per-role counters: ~100 cycles per MMA inside the issue loop, i.e. the pipe's own
        // pace, plus ~410 cycles per stage outside it).  Hence: stage slot / phase tracked incrementally (no div/mod),
        // operand descriptors advanced from two bases built once, cycle counters only when debugging.
        const uint32_t idesc = instr_desc(BM, nstack, 0, 0);
        const bool dbg = p.dbg != nullptr && blockIdx.x == 0;
        const long long t_begin = clock64();
        long long t_w = 0, t_full = 0, t_tmem = 0, t_issue = 0, c0 = 0;
        if (!STREAM) {
            mbar_wait(smem_u32(&bars[2 * S + 4]), 0);
            t_w = clock64() - t_begin;
            tc_fence_after();
            if (my_items == 0 && p.chunks > 1) mbar_wait(smem_u32(&bars[2 * S + 5]), 0);   // in-flight TMA must land
        }
        const uint64_t ad_base = smem_desc(smem_u32(smem_a), 16, 1024);
        const uint64_t bd_base = smem_desc(STREAM ? smem_u32(smem_a) + KX_BOX_BYTES : smem_u32(smem), 16, 1024);
        const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[S]);
        const uint32_t tfull0 = smem_u32(&bars[2 * S]), tempty0 = smem_u32(&bars[2 * S + 2]);
        const int last_ksteps = (p.cin - (p.chunks - 1) * KCH + 7) / 8;
        const int last_ch = p.chunks - 1;
        int s = 0, it = 0;
        uint32_t ph = 0, s_off = 0;                 // phase of stage slot s; byte offset of slot s
        for (int t = 0; t < my_items; ++t) {
            const int buf = t & 1;
            if (dbg) c0 = clock64();
            mbar_wait(tempty0 + 8u * (uint32_t)buf, (((uint32_t)t >> 1) & 1u) ^ 1u);   // epilogue drained this buffer
            if (dbg) t_tmem += clock64() - c0;
            tc_fence_after();
            const uint32_t d_addr = tmem_base + (uint32_t)(buf * nstack);
            for (int ch = 0; ch <= last_ch; ++ch, ++it) {
                if (dbg) c0 = clock64();
                mbar_wait(full0 + 8u * (uint32_t)s, ph);
                if (dbg) t_full += clock64() - c0;
                if (!STREAM && t == 0 && ch == 1) {
                    if (dbg) c0 = clock64();
                    mbar_wait(smem_u32(&bars[2 * S + 5]), 0);      // second half of the filter bank
                    if (dbg) t_w += clock64() - c0;
                }
                tc_fence_after();
                if (dbg) c0 = clock64();
                if (elect_one()) {
                    const uint64_t ad0 = desc_advance(ad_base, s_off);
                    const uint64_t bd0 = desc_advance(bd_base, STREAM ? s_off : (uint32_t)ch * b_chunk);
                    if (ch < last_ch || last_ksteps == 4) { MI_KXS_MMAS(4) }
                    else if (last_ksteps == 3) { MI_KXS_MMAS(3) }
                    else if (last_ksteps == 2) { MI_KXS_MMAS(2) }
                    else { MI_KXS_MMAS(1) }
                    umma_commit(empty0 + 8u * (uint32_t)s);
                    if (ch == last_ch) umma_commit(tfull0 + 8u * (uint32_t)buf);
                }
                __syncwarp();
                if (dbg) t_issue += clock64() - c0;
                if (++s == S) { s = 0; s_off = 0; ph ^= 1u; } else { s_off += stage_bytes; }
            }
        }
        if (dbg && lane == 0) {
            p.dbg[2] = (unsigned long long)t_w; p.dbg[3] = (unsigned long long)t_full;
            p.dbg[4] = (unsigned long long)t_tmem; p.dbg[5] = (unsigned long long)(clock64() - t_begin);
            p.dbg[11] = (unsigned long long)t_issue; p.dbg[12] = (unsigned long long)it;
        }
    } else if (warp < 10) {
        const int q = warp & 3;               // TMEM lane quarter this warp may read
        const int half = (warp - 2) >> 2;     // which 16-column half of every 32-column chunk it owns
        const int r = q * 32 + lane;
        const int row_i = r >> 4, col_i = r & 15;
        const bool col_ok = (col_i >= 1) && (col_i <= KX_OW);
        long long e_wait = 0, e_ld = 0, e_st = 0, e_b1 = 0;
        const long long e_begin = clock64();
        EpiArgs ea;
        ea.cout = p.cout; ea.act = p.act; ea.mask_act = p.mask_act; ea.accumulate = p.accumulate; ea.rnd = p.rnd;
        ea.slope = p.slope; ea.mask_slope = p.mask_slope;
        ea.cout_store = p.cout;   // lanes past cout are never written: they may belong to the next concat slice
        // staged pixel row of this thread: slot = row * 14 + col - 1, 128 bytes per 32-channel box, 16-byte chunks XOR-ed
        // with slot & 7 (the 128B swizzle the TMA store undoes); a quarter warp holds eight consecutive slots, so its
        // 16-byte accesses fall into eight different bank groups
        const int slot = row_i * KX_OW + col_i - 1;
        const uint32_t stg = smem_u32(staging);
        const uint32_t my_row = stg + (uint32_t)slot * ROW_BYTES;
        const uint32_t swz = (uint32_t)(slot & 7);
        // The tensor maps of y / the operand cover the channels below cout & ~3: TMA clips the innermost dimension at
        // 16-byte granularity (a 18-channel slice of a 20-channel buffer had its neighbours 18..19 overwritten), so a
        // ragged tail of 1-3 channels is stored by the thread that holds it, with plain scalar accesses.
        // (When the caller owns the pad lanes of the row -- mi_set_pad_lanes_scratch -- c_tma is cout rounded UP to 4 and
        // the pad lanes are stored as zeros.)
        const int c_tma = p.c_tma, c_tail = c_tma < p.cout ? p.cout - c_tma : 0;
        const uint32_t op_bar = smem_u32(&bars[2 * S + 6]);
        auto item_coords = [&](int t, int& co0, int& x0, int& y0, int& img) {
            const int item = (int)blockIdx.x + t * (int)gridDim.x;
            const int nt = STREAM ? item % p.n_tiles : 0;
            co0 = nt * p.bn;
            int tile = STREAM ? item / p.n_tiles : item;
            const int tx_i = tile % p.tiles_x; tile /= p.tiles_x;
            const int ty_i = tile % p.tiles_y; tile /= p.tiles_y;
            img = tile; x0 = tx_i * KX_OW; y0 = ty_i * KX_H;
        };
        for (int t = 0; t < my_items; ++t) {
            const int buf = t & 1;
            const uint32_t use = (uint32_t)(t >> 1);
            int co0, x0, y0, img;
            item_coords(t, co0, x0, y0, img);
            long long c0 = clock64();
            mbar_wait(smem_u32(&bars[2 * S + buf]), use & 1u);
            e_wait += clock64() - c0;
            tc_fence_after();
            // up to two 16-channel pieces per warp (bn <= 64): read and combine them, release the buffer, then store
            float acc[2][16];
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * nstack);
            const int npiece = (p.bn > 32 && co0 + 32 < p.cout) ? 2 : 1;       // CTA-uniform (box granularity)
            long long c1 = clock64();
#pragma unroll
            for (int pc = 0; pc < 2; ++pc) {
                if (pc < npiece) {
                    const int c0i = 16 * half + 32 * pc;
                    uint32_t d0[16], d1[16], d2[16];
                    tmem_ld16_nowait(t_row + (uint32_t)c0i, d0);
                    tmem_ld16_nowait(t_row + (uint32_t)(p.bn + c0i), d1);
                    tmem_ld16_nowait(t_row + (uint32_t)(2 * p.bn + c0i), d2);
                    tmem_ld_wait();
                    const float* sb = sbias + co0 + c0i;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float left = __shfl_up_sync(0xffffffffu, __uint_as_float(d0[j]), 1);
                        const float right = __shfl_down_sync(0xffffffffu, __uint_as_float(d2[j]), 1);
                        acc[pc][j] = ((left + __uint_as_float(d1[j])) + right) + sb[j];
                    }
                }
            }
            e_ld += clock64() - c1;
            tc_fence_before();
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars[2 * S + 2 + buf])) : "memory");
            c1 = clock64();
            if (ea.act == MI_ACT_RELU) {
#pragma unroll
                for (int pc = 0; pc < 2; ++pc)
#pragma unroll
                    for (int j = 0; j < 16; ++j) acc[pc][j] = fmaxf(acc[pc][j], 0.f);
            } else if (ea.act == MI_ACT_LEAKY) {
#pragma unroll
                for (int pc = 0; pc < 2; ++pc)
#pragma unroll
                    for (int j = 0; j < 16; ++j) acc[pc][j] = acc[pc][j] > 0.f ? acc[pc][j] : acc[pc][j] * ea.slope;
            } else if (ea.act != MI_ACT_NONE) {
#pragma unroll
                for (int pc = 0; pc < 2; ++pc)
#pragma unroll
                    for (int j = 0; j < 16; ++j) acc[pc][j] = kxs_act_rare(acc[pc][j], ea.act, ea.slope);
            }
            if (EPI == KXS_EPI_GENERIC) {
                const int oy = y0 + row_i, ox = x0 + col_i - 1;
                const bool pix_ok = col_ok && (oy < p.h) && (ox < p.w);
                const long long pix = ((long long)img * p.h + oy) * p.w + ox;
                const float* mrow = p.mask_y ? p.mask_y + pix * p.ldmask : nullptr;
                float* yrow = p.y + pix * p.ldy;
                for (int pc = 0; pc < npiece; ++pc) {
                    const int co = co0 + 16 * half + 32 * pc;
                    if (!pix_ok) continue;
                    for (int j = 0; j < 16; ++j) {
                        if (co + j >= p.cout) break;
                        float val = pc ? acc[1][j] : acc[0][j];
                        if (mrow) val *= mi_act_grad(__ldg(mrow + co + j), ea.mask_act, ea.mask_slope);
                        if (ea.accumulate) val += yrow[co + j];
                        yrow[co + j] = ea.rnd ? mi_rn_tf32(val) : val;
                    }
                }
            } else {
              // one round per item -- or, with a single staging box (p.one_box, plain mode), one round per 32-channel box
              const int rounds = (EPI == KXS_EPI_PLAIN && p.one_box) ? npiece : 1;
              // Cout <= 32: an item fills ONE box, so items alternate between the two boxes of the staging tile and the
              // store of item t overlaps the staging of item t + 1 (these layers are the cheapest per tile -- 12 MMAs
              // of N = 96 -- and were paced by the store round trip, per-role counters)
              const uint32_t alt_off = (p.bn <= 32 && !p.one_box) ? (uint32_t)(t & 1) * KX_OUT_BOX : 0u;
              for (int rd = 0; rd < rounds; ++rd) {
                const long long cb = clock64();
                if (EPI == KXS_EPI_OPERAND) mbar_wait(op_bar, (uint32_t)t & 1u);   // operand landed => staging is ours
                else kxs_bar_sync(1);                                               // previous store has left the tile
                e_b1 += clock64() - cb;
                if (col_ok) {
#pragma unroll
                    for (int pc = 0; pc < 2; ++pc) {
                        if (pc < npiece && (rounds == 1 || pc == rd)) {
#pragma unroll
                            for (int g = 0; g < 4; ++g) {
                                const uint32_t a = my_row + (rounds == 1 ? (uint32_t)pc * KX_OUT_BOX : 0u) + alt_off +
                                                   ((((uint32_t)(4 * half + g)) ^ swz) << 4);
                                float4 v = make_float4(acc[pc][4 * g], acc[pc][4 * g + 1], acc[pc][4 * g + 2], acc[pc][4 * g + 3]);
                                if (EPI == KXS_EPI_OPERAND) {
                                    float4 o;
                                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                                 : "=f"(o.x), "=f"(o.y), "=f"(o.z), "=f"(o.w) : "r"(a));
                                    if (ea.accumulate) {
                                        v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
                                    } else if (ea.mask_act == MI_ACT_RELU) {
                                        v.x = o.x > 0.f ? v.x : 0.f; v.y = o.y > 0.f ? v.y : 0.f;
                                        v.z = o.z > 0.f ? v.z : 0.f; v.w = o.w > 0.f ? v.w : 0.f;
                                    } else {
                                        v.x *= kxs_act_grad_rare(o.x, ea.mask_act, ea.mask_slope);
                                        v.y *= kxs_act_grad_rare(o.y, ea.mask_act, ea.mask_slope);
                                        v.z *= kxs_act_grad_rare(o.z, ea.mask_act, ea.mask_slope);
                                        v.w *= kxs_act_grad_rare(o.w, ea.mask_act, ea.mask_slope);
                                    }
                                }
                                if (ea.rnd) v = mi_rn_tf32(v);
                                const int cg = co0 + 32 * pc + 16 * half + 4 * g;
                                if (cg + 4 > p.cout) {              // pad lanes (and the channels past them) hold zeros
                                    if (cg + 1 > p.cout) v.x = 0.f;
                                    if (cg + 2 > p.cout) v.y = 0.f;
                                    if (cg + 3 > p.cout) v.z = 0.f;
                                    v.w = 0.f;
                                }
                                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};"
                                             ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
                            }
                        }
                    }
                }
                if (rd == 0 && c_tail && c_tma >= co0 && c_tma < co0 + p.bn && half == (((c_tma - co0) >> 4) & 1)) {
                    const int oy = y0 + row_i, ox = x0 + col_i - 1;
                    if (col_ok && oy < p.h && ox < p.w) {
                        const long long pix = ((long long)img * p.h + oy) * p.w + ox;
                        const int pc_t = (c_tma - co0) >> 5, g_t = ((c_tma - co0) >> 2) & 3;
                        float tv[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                        for (int pc = 0; pc < 2; ++pc)
#pragma unroll
                            for (int g = 0; g < 4; ++g)
                                if (pc == pc_t && g == g_t) {
                                    tv[0] = acc[pc][4 * g]; tv[1] = acc[pc][4 * g + 1]; tv[2] = acc[pc][4 * g + 2];
                                }
                        for (int qq = 0; qq < c_tail; ++qq) {
                            float val = tv[qq];
                            if (p.mask_y)
                                val *= kxs_act_grad_rare(__ldg(p.mask_y + pix * p.ldmask + c_tma + qq), ea.mask_act, ea.mask_slope);
                            float* dst = p.y + pix * p.ldy + c_tma + qq;
                            if (ea.accumulate) val += *dst;
                            *dst = ea.rnd ? mi_rn_tf32(val) : val;
                        }
                    }
                }
                fence_async_smem();
                kxs_bar_arrive(2);                  // staged: the store warp takes it from here
              }
            }
            e_st += clock64() - c1;
        }
        if (p.dbg && blockIdx.x == 0 && threadIdx.x == 64) {
            p.dbg[6] = (unsigned long long)e_wait; p.dbg[7] = (unsigned long long)e_ld;
            p.dbg[8] = (unsigned long long)e_st; p.dbg[9] = (unsigned long long)(clock64() - e_begin);
            p.dbg[10] = (unsigned long long)my_items;
            p.dbg[14] = (unsigned long long)e_b1;
        }
    } else if (EPI != KXS_EPI_GENERIC) {
        // store warp: per item, wait for the staged tile, issue one bulk store per 32-channel box, wait until TMA has
        // read the tile, then hand it back -- as the landing zone of the next item's operand (mask / previous y), or
        // plainly free
        const uint32_t stg = smem_u32(staging);
        const uint32_t op_bar = smem_u32(&bars[2 * S + 6]);
        const int c_tma = p.c_tma;
        long long e_wr = 0;
        auto item_coords = [&](int t, int& co0, int& x0, int& y0, int& img) {
            const int item = (int)blockIdx.x + t * (int)gridDim.x;
            const int nt = STREAM ? item % p.n_tiles : 0;
            co0 = nt * p.bn;
            int tile = STREAM ? item / p.n_tiles : item;
            const int tx_i = tile % p.tiles_x; tile /= p.tiles_x;
            const int ty_i = tile % p.tiles_y; tile /= p.tiles_y;
            img = tile; x0 = tx_i * KX_OW; y0 = ty_i * KX_H;
        };
        const bool alt = p.bn <= 32 && !p.one_box;          // items alternate between the two staging boxes
        auto load_operand = [&](int t) {          // one lane: mask / previous-y boxes of item t into the staging tile
            int co0, x0, y0, img;
            item_coords(t, co0, x0, y0, img);
            const int nbox = (p.bn > 32 && co0 + 32 < c_tma) ? 2 : 1;
            const uint32_t dst = stg + (alt ? (uint32_t)(t & 1) * KX_OUT_BOX : 0u);
            mbar_expect_tx(op_bar, (uint32_t)nbox * KX_OUT_BOX);
            for (int b = 0; b < nbox; ++b)
                tma_load_4d(dst + (uint32_t)b * KX_OUT_BOX, &map_op, op_bar, co0 + 32 * b, x0, y0, img);
        };
        if (EPI == KXS_EPI_OPERAND) {
            if (my_items > 0 && elect_one()) load_operand(0);
            __syncwarp();
        } else {
            kxs_bar_arrive(1);                       // the staging tile starts out free
        }
        for (int t = 0; t < my_items; ++t) {
            int co0, x0, y0, img;
            item_coords(t, co0, x0, y0, img);
            const int npiece = (p.bn > 32 && co0 + 32 < p.cout) ? 2 : 1;       // as the epilogue warps count them
            const int rounds = (EPI == KXS_EPI_PLAIN && p.one_box) ? npiece : 1;
            for (int rd = 0; rd < rounds; ++rd) {
                const bool last = (t + 1 == my_items) && (rd + 1 == rounds);
                kxs_bar_sync(2);                     // all eight epilogue warps have staged item t (round rd)
                if (elect_one()) {
                    if (alt) {
                        if (co0 < c_tma) tma_store_4d(&map_y, stg + (uint32_t)(t & 1) * KX_OUT_BOX, co0, x0, y0, img);
                    } else if (rounds == 1) {
                        for (int b = 0; b < 2; ++b)
                            if (32 * b < p.bn && co0 + 32 * b < c_tma)
                                tma_store_4d(&map_y, stg + (uint32_t)b * KX_OUT_BOX, co0 + 32 * b, x0, y0, img);
                    } else if (co0 + 32 * rd < c_tma) {
                        tma_store_4d(&map_y, stg, co0 + 32 * rd, x0, y0, img);
                    }
                    tma_store_commit();
                    const long long cw = clock64();
                    if (alt) tma_store_wait_read_but_one();     // the OTHER box (item t - 1) has been read: it is free
                    else tma_store_wait_read();
                    e_wr += clock64() - cw;
                    if (EPI == KXS_EPI_OPERAND && t + 1 < my_items) load_operand(t + 1);
                    if (last) tma_store_wait_all();
                    if (p.dbg && blockIdx.x == 0 && last) p.dbg[13] = (unsigned long long)e_wr;
                }
                __syncwarp();
                if (EPI == KXS_EPI_PLAIN && !last) kxs_bar_arrive(1);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}
