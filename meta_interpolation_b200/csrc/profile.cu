// Optional per-launch timing of the hot kernels with CUDA events on the launching stream
// (bench.py's roofline section).  Disabled by default; never used under graph capture.
#include <vector>

#include "mi_common.cuh"

namespace {
struct Rec {
    int tag;
    double flops, bytes;
    cudaEvent_t start, stop;
};
bool g_enabled = false;
std::vector<Rec> g_recs;
std::vector<cudaEvent_t> g_pool;
size_t g_pool_used = 0;

cudaEvent_t take_event() {
    if (g_pool_used == g_pool.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        g_pool.push_back(e);
    }
    return g_pool[g_pool_used++];
}
}  // namespace

void mi_prof_begin(int tag, double flops, double bytes, cudaStream_t s) {
    if (!g_enabled) return;
    Rec r;
    r.tag = tag; r.flops = flops; r.bytes = bytes;
    r.start = take_event();
    r.stop = take_event();
    cudaEventRecord(r.start, s);
    g_recs.push_back(r);
}

void mi_prof_end(cudaStream_t s) {
    if (!g_enabled || g_recs.empty()) return;
    cudaEventRecord(g_recs.back().stop, s);
}

extern "C" {

int mi_prof_enable(int on) {
    g_enabled = on != 0;
    g_recs.clear();
    g_pool_used = 0;
    return MI_OK;
}

// out[0]=launches, out[1]=total ms, out[2]=algorithmic flops, out[3]=algorithmic bytes for `tag`
int mi_prof_summary(int tag, double* out) {
    if (!out) return MI_ERR_BAD_ARG;
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return (int)e;
    out[0] = out[1] = out[2] = out[3] = 0.0;
    for (const Rec& r : g_recs) {
        if (r.tag != tag) continue;
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.start, r.stop) != cudaSuccess) { cudaGetLastError(); continue; }
        out[0] += 1.0; out[1] += ms; out[2] += r.flops; out[3] += r.bytes;
    }
    return MI_OK;
}

}  // extern "C"
