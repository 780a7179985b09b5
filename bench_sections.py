"""bench.py helpers: the roofline section (per-launch CUDA-event timing of the dominant kernel), the
bounded CPU baseline and the ``--impl reference`` arm (the pinned oracle port on the host cores)."""
import json
import os
import time

import torch

import bench

ROOT = os.path.dirname(os.path.abspath(__file__))


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            d = json.load(f)
        return dict(hbm=d.get("hbm_gbs", 6650.0), bf16_burst=d.get("bf16_tflops", 1590.0),
                    bf16_sustained=d.get("bf16_tflops_sustained", 1400.0), source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


def roofline_section(system):
    """One extra instrumented meta-iteration (1 task, eager launches, CUDA events recorded around every hot
    kernel on the launch stream) after the timed region.  The dominant kernel class is the one with the
    largest share of the instrumented step; `achieved` = its algorithmic FLOPs (2*MACs of the true conv
    shape) / its summed event time."""
    ops = system.ops
    frames = [f.cuda() for f in bench.synthetic_septuplets(1, 999)]
    saved = system.use_cuda_graphs
    system.use_cuda_graphs = False
    fast = system.fast_path()
    fast_saved = fast.use_graphs
    fast.use_graphs = False
    try:
        system.run_train_iter(frames, epoch=0)        # untimed (allocator warm-up for batch 1 shapes)
        torch.cuda.synchronize()
        ops.prof_enable(True)
        t0 = time.perf_counter()
        system.run_train_iter(frames, epoch=0)
        summ = ops.prof_summary()
        wall_ms = (time.perf_counter() - t0) * 1e3
        ops.prof_enable(False)
    finally:
        system.use_cuda_graphs = saved
        fast.use_graphs = fast_saved
    peaks = measured_peaks()
    total_ms = sum(v["ms"] for v in summ.values())
    conv_tags = ("fprop_tc_halo", "fprop_tc_halo_stream", "fprop_tc", "wgrad_tc_kx", "wgrad_tc", "fprop_simt",
                 "wgrad_simt")
    dom = max(conv_tags, key=lambda t: summ[t]["ms"])
    d = summ[dom]
    achieved = d["flops"] / (d["ms"] * 1e-3) / 1e12 if d["ms"] > 0 else 0.0
    peak = peaks["bf16_sustained"]
    breakdown = {k: {"launches": v["launches"], "ms": round(v["ms"], 3),
                     "tflops": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 2) if v["ms"] > 0 else 0.0,
                     "gbs": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["ms"] > 0 else 0.0}
                 for k, v in summ.items() if v["launches"]}
    return {
        "bound": "tensor", "kernel": KERNEL_NAMES.get(dom, dom), "achieved": round(achieved, 2), "peak": peak,
        "unit": "TFLOP/s", "frac": round(achieved / peak, 4),
        # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel on its heaviest
        # layer (51->51 on the 258x450 region of interest, N=2): see profiles/README.md
        "traffic": NCU_TRAFFIC_BYTES.get(dom),
        "traffic_launch": NCU_TRAFFIC_LAUNCH.get(dom),
        "peak_source": peaks["source"] + "; dense bf16 sustained (kernel timed inside a long step); the kernel "
                       "computes in TF32, whose tensor-pipe ceiling is half the bf16 one (frac 0.5 = TF32 peak)",
        "share_of_instrumented_step": round(d["ms"] / total_ms, 3) if total_ms > 0 else None,
        "avg_launch_us": round(d["ms"] * 1e3 / max(d["launches"], 1), 2),
        "algorithmic_gflop_per_launch": round(d["flops"] / max(d["launches"], 1) / 1e9, 3),
        "instrumented_step_ms": round(wall_ms, 2), "per_kernel": breakdown,
    }


KERNEL_NAMES = {"fprop_tc_halo": "conv_fprop_tc_halo_kernel", "fprop_tc_halo_stream": "conv_fprop_tc_halo_stream_kernel",
                "fprop_tc": "conv_fprop_tc_kernel", "wgrad_tc_kx": "conv_wgrad_tc_kx_kernel",
                "wgrad_tc": "conv_wgrad_tc_kernel", "fprop_simt": "conv_fprop_simt_kernel",
                "wgrad_simt": "conv_wgrad_simt_kernel"}
# filled from the committed ncu captures (profiles/); None = not captured this round.  The capture is of ONE launch
# (named in NCU_TRAFFIC_LAUNCH, with its algorithmic bytes), while `achieved` averages all launches of the kernel.
NCU_TRAFFIC_LAUNCH = {"fprop_tc_halo": "51->51 3x3 on the 258x450 region of interest, N=2: 10.9 GFLOP, 94.7 MB "
                                       "algorithmic (profiles/r01c_ncu_conv_fprop_halo_51x51_258x450.txt)"}
NCU_TRAFFIC_BYTES = {"fprop_tc_halo": 48479488 + 6179072}   # dram read + write of that launch


# --------------------------------------------------------------------------------------------- CPU legs
def _oracle_system(batch):
    from oracle import backbones as bb, maml
    return maml.OracleSystem('sepconv', bb.seeded_params('sepconv', 12345), optimizer='SGD', num_steps=bench.K_INNER,
                             inner_lr=1e-5, outer_lr=1e-5, loss='1*L1')


def _cpu_pass(system, frames):
    """A bounded sample of the workload on the host: one support step (2 support triplets: forward, L1,
    backward, LSLR-SGD update) + nothing else.  A task costs K support steps + 1 query pass = (2K+1)/2 = 5.5
    of these in forward/backward work."""
    from collections import OrderedDict
    fast = OrderedDict(system.params)
    sl = system._support_loss(frames, 0, fast, ((0, 2, 4), (2, 4, 6)))
    fast, _ = system.inner_update(sl, fast, {}, 0)
    return fast


def cpu_baseline_section():
    """The oracle port (oracle/maml.py, ATen CPU ops, all host threads) on one support step of one
    256x448 task; tasks/s extrapolated by the pass count (2K+1 passes per task, 2 per sample)."""
    threads = torch.get_num_threads()
    system = _oracle_system(1)
    frames = bench.synthetic_septuplets(1, 100)
    t0 = time.perf_counter()
    _cpu_pass(system, frames)
    dt = time.perf_counter() - t0
    per_task = dt * (2 * bench.K_INNER + 1) / 2.0
    return {"value": round(1.0 / per_task, 5), "unit": "tasks/s", "cores": threads, "kind": "port",
            "sample": "1 support step (2 support triplets fwd+bwd+update) of one 256x448 task = 2 of the 11 passes of "
                      "a task; %.1f s measured, task time extrapolated x5.5" % dt}


def cpu_reference_run(steps, warmup, world):
    """--impl reference: the reference's algorithm on the host cores (pinned oracle port; the reference itself
    is unpackaged Python that cannot travel to the GPU box).  Each step is the bounded sample above."""
    threads = torch.get_num_threads()
    system = _oracle_system(1)
    frames = bench.synthetic_septuplets(1, 100)
    for _ in range(min(warmup, 1)):
        _cpu_pass(system, frames)
    t0 = time.perf_counter()
    for _ in range(steps):
        _cpu_pass(system, frames)
    dt = (time.perf_counter() - t0) / steps
    per_task = dt * (2 * bench.K_INNER + 1) / 2.0
    value = 1.0 / per_task
    return {
        "impl": "reference", "metric": bench.METRIC, "value": round(value, 5), "unit": "tasks/s", "n_gpus": world,
        "steps": steps, "warmup": min(warmup, 1), "ms_per_step": round(dt * 1e3, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": bench.WORKLOAD},
        "cpu_baseline": {"value": round(value, 5), "unit": "tasks/s", "cores": threads, "kind": "port",
                         "sample": "each step = 1 support step of one 256x448 task (2 of its 11 passes), x5.5"},
        "e2e": {"value": round(value, 5), "unit": "tasks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
