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
        # the same instrumented iteration at the SM share of the timed run (8 tasks in flight: a persistent launch
        # takes a quarter of the SMs and four of them run side by side)
        fast.force_sm_shares = 4
        system.run_train_iter(frames, epoch=0)
        torch.cuda.synchronize()
        ops.prof_enable(True)
        system.run_train_iter(frames, epoch=0)
        summ_share = ops.prof_summary()
        ops.prof_enable(False)
        share_ctas = ops.set_sm_budget(0) // 4
    finally:
        fast.force_sm_shares = None
        system.use_cuda_graphs = saved
        fast.use_graphs = fast_saved
    # device time of the same one-task step replayed as CUDA graphs on ONE stream (no launch gaps, no lane overlap):
    # the denominator of `share_of_step`, which therefore counts every kernel of the step, tagged or not
    step_ms = None
    if fast_saved:
        lanes_saved = fast.n_lanes
        fast.n_lanes = 1
        try:
            for _ in range(3):
                system.run_train_iter(frames, epoch=0)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            system.run_train_iter(frames, epoch=0)
            e1.record()
            torch.cuda.synchronize()
            step_ms = e0.elapsed_time(e1)
        finally:
            fast.n_lanes = lanes_saved
    peaks = measured_peaks()
    total_ms = sum(v["ms"] for v in summ.values())
    conv_tags = ("fprop_tc_kxs", "fprop_tc_halo", "fprop_tc_halo_stream", "fprop_tc", "wgrad_tc_kx", "wgrad_tc",
                 "fprop_simt", "wgrad_simt")
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
        "share_of_tagged_kernels": round(d["ms"] / total_ms, 3) if total_ms > 0 else None,
        "share_of_step": round(d["ms"] / step_ms, 3) if step_ms else None,
        "one_task_step_ms_single_stream_graphs": round(step_ms, 3) if step_ms else None,
        "avg_launch_us": round(d["ms"] * 1e3 / max(d["launches"], 1), 2),
        "algorithmic_gflop_per_launch": round(d["flops"] / max(d["launches"], 1) / 1e9, 3),
        "instrumented_step_ms": round(wall_ms, 2), "per_kernel": breakdown,
        # the timed run does not launch this kernel on the whole chip: with 8 tasks in flight a launch gets a quarter
        # of the SMs (fastpath.py).  Same per-launch event timing at that share; `frac_of_share` = achieved / (peak x
        # CTAs / SMs), the fraction of what the launch's own SMs can do (the headline `frac` above stays whole-chip)
        "at_sm_share": _share_view(summ_share[dom], share_ctas, ops.set_sm_budget(0), peak),
    }


def _share_view(d, ctas, sms, peak):
    if not d["launches"] or d["ms"] <= 0:
        return None
    achieved = d["flops"] / (d["ms"] * 1e-3) / 1e12
    return {"ctas_per_launch": ctas, "sms": sms, "achieved": round(achieved, 2), "unit": "TFLOP/s",
            "avg_launch_us": round(d["ms"] * 1e3 / d["launches"], 2),
            "frac_of_share": round(achieved / (peak * ctas / sms), 4)}


KERNEL_NAMES = {"fprop_tc_kxs": "conv_fprop_tc_kxs_kernel", "fprop_tc_halo": "conv_fprop_tc_halo_kernel", "fprop_tc_halo_stream": "conv_fprop_tc_halo_stream_kernel",
                "fprop_tc": "conv_fprop_tc_kernel", "wgrad_tc_kx": "conv_wgrad_tc_kx_kernel",
                "wgrad_tc": "conv_wgrad_tc_kernel", "fprop_simt": "conv_fprop_simt_kernel",
                "wgrad_simt": "conv_wgrad_simt_kernel"}
# filled from the committed ncu captures (profiles/); None = not captured this round.  The capture is of ONE launch
# (named in NCU_TRAFFIC_LAUNCH, with its algorithmic bytes), while `achieved` averages all launches of the kernel.
NCU_TRAFFIC_LAUNCH = {"fprop_tc_kxs": "51->51 3x3 on the 258x450 region of interest, N=2: 10.9 GFLOP, 94.7 MB "
                                      "algorithmic (profiles/r02b_ncu_kxs_51x51_258x450.txt)",
                      "fprop_tc_halo": "51->51 3x3 on the 258x450 region of interest, N=2: 10.9 GFLOP, 94.7 MB "
                                       "algorithmic (profiles/r01c_ncu_conv_fprop_halo_51x51_258x450.txt)"}
# dram read + write of that launch (the output of the layer mostly stays in L2, hence less than the algorithmic bytes)
NCU_TRAFFIC_BYTES = {"fprop_tc_kxs": 48465920 + 4976896, "fprop_tc_halo": 48479488 + 6179072}


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


def _all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm runs on rank 0 alone and may use the box."""
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        pass
    torch.set_num_threads(n)
    return torch.get_num_threads()


def cpu_reference_run(steps, warmup, world):
    """--impl reference: the reference's algorithm on the host cores (pinned oracle port; SepConv has no CPU path in
    the reference itself, sepconv.py:293-294).  Each step is the bounded sample above."""
    threads = _all_host_threads()
    system = _oracle_system(1)
    frames = bench.synthetic_septuplets(1, 100)
    for _ in range(warmup):
        _cpu_pass(system, frames)
    t0 = time.perf_counter()
    for _ in range(steps):
        _cpu_pass(system, frames)
    dt = (time.perf_counter() - t0) / steps
    per_task = dt * (2 * bench.K_INNER + 1) / 2.0
    value = 1.0 / per_task
    return {
        "impl": "reference", "metric": bench.METRIC, "value": round(value, 5), "unit": "tasks/s", "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": round(dt * 1e3, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench.bench_config(world),
        "cpu_baseline": {"value": round(value, 5), "unit": "tasks/s", "cores": threads, "kind": "port",
                         "sample": "each step = 1 support step of one 256x448 task (2 of its 11 passes); "
                                   "tasks/s = 1 / (5.5 x step time); rank 0 only, all host threads"},
        "e2e": {"value": round(value, 5), "unit": "tasks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


# --------------------------------------------------------------------------------------------- other configs
OTHER_CONFIGS = {
    # BASELINE.json configs[0] (the reference's CPU plumbing case, here on the GPU: one task per GPU) and configs[2..4]:
    # (flags, (H, W), global meta-batch, scaling, init gain)
    "C1 voxelflow K=1 128x128": (dict(model="voxelflow", loss="1*MSE", optimizer="SGD",
                                      number_of_training_steps_per_iter=1), (128, 128), 1, "weak", None),
    "C3 superslomo Meta-SGD K=5 256x448": (dict(model="superslomo", loss="1*L1", optimizer="SGD", metasgd=True,
                                                number_of_training_steps_per_iter=5), (256, 448), 4, "weak", None),
    "C4 cain L2F K=3 512x512": (dict(model="cain", loss="1*L1", optimizer="SGD", attenuate=True,
                                     number_of_training_steps_per_iter=3), (512, 512), 4, "weak", 0.4),
    "C5 rrin MAML++ MSL K=5 256x448": (dict(model="rrin", loss="1*L1", optimizer="SGD",
                                            number_of_training_steps_per_iter=5,
                                            learnable_per_layer_per_step_inner_loop_learning_rate=True,
                                            use_multi_step_loss_optimization=True, multi_step_loss_num_epochs=1),
                                       (256, 448), 64, "strong", None),
}


def _normalise(frames, model):
    if model == "superslomo":
        mean = torch.tensor([0.429, 0.431, 0.397]).view(1, 3, 1, 1)
        return [f - mean for f in frames]
    if model == "voxelflow":
        return [(f * 255 - 127.5) / 127.5 for f in frames]
    return frames


def other_configs_section(rank, world, timed):
    """BASELINE configs[2..4] after the headline's timed region: 3 warm-up + 3 timed `run_train_iter` each, frames
    resident in HBM, device time (max over ranks).  C3 / C4 keep 4 tasks per GPU (weak scaling: 32 tasks on 8 GPUs,
    16 on 4); C5 is the strong-scaling sweep, 64 tasks split over the N ranks."""
    import gc
    from meta_interpolation_b200.meta_learning_system import SceneAdaptiveInterpolation
    out = {}
    for name, (over, (h, w), tasks, scaling, gain) in OTHER_CONFIGS.items():
        batch = tasks * world if scaling == "weak" else tasks
        if batch % world:
            out[name] = {"skipped": "%d tasks do not split over %d ranks" % (batch, world)}
            continue
        a = bench.make_args(batch)
        for k, v in over.items():
            setattr(a, k, v)
        a.number_of_evaluation_steps_per_iter = a.number_of_training_steps_per_iter
        system = SceneAdaptiveInterpolation(a)
        if gain is not None:   # cain's default init explodes through 125 stacked convs: seeded init x0.4
            with torch.no_grad():
                for p in system.net.parameters():
                    if p.dim() == 4:
                        p.mul_(gain)
        per = batch // world
        local = bench.synthetic_septuplets(per, 321 + rank, h, w)
        frames = [torch.zeros(batch, 3, h, w, device="cuda") for _ in range(7)]
        for t in range(7):      # only this rank's tasks are ever read
            frames[t][rank * per:(rank + 1) * per].copy_(_normalise(local, over["model"])[t])
        for _ in range(3):
            system.run_train_iter(frames, epoch=0)
        steps = 3
        ms = timed(lambda i: system.run_train_iter(frames, epoch=0), steps)
        out[name] = {"value": round(batch * steps / (ms / 1e3), 3), "unit": "tasks/s", "global_batch": batch,
                     "tasks_per_gpu": per, "scaling": scaling, "ms_per_step": round(ms / steps, 2), "steps": steps,
                     "warmup": 3, "graph_path": bool(system.fast_path_supported())}
        del system, frames
        gc.collect()
        torch.cuda.empty_cache()
    return out


# --------------------------------------------------------------------------------------------- GPU reference
def gpu_reference_section(our_e2e):
    """north_star's denominator: the UNMODIFIED reference (staged copy under baseline/_ref, its own cupy kernel
    strings through NVRTC, cuDNN with its default allow_tf32=True) timed on the same GPU right after our arm, same
    workload (8 tasks per step, frames resident on the device -- the reference's `run_train_iter` takes device
    tensors).  Runs in a subprocess with one visible GPU; `ratio` = our end-to-end tasks/s / its tasks/s."""
    import subprocess
    import sys
    script = os.path.join(ROOT, "baseline", "reference_gpu.py")
    if not os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "meta_learning_system.py")):
        return {"unavailable": "baseline/_ref (staged copy of the reference) is not present"}
    env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", "0").split(",")[0])
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT", "OMP_NUM_THREADS"):
        env.pop(k, None)
    out = {}
    for tf32 in (True, False):
        cmd = [sys.executable, script, "--model", "sepconv", "--batch", str(bench.TASKS_PER_GPU), "--steps", "5",
               "--warmup", "3"] + ([] if tf32 else ["--no-tf32"])
        try:
            res = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
            line = json.loads(res.stdout.strip().splitlines()[-1]) if res.returncode == 0 else \
                {"error": res.stderr.strip().splitlines()[-1][:300] if res.stderr.strip() else "rc %d" % res.returncode}
        except Exception as e:      # noqa: BLE001 -- a failed reference run must not take the bench line with it
            line = {"error": repr(e)[:300]}
        out["allow_tf32" if tf32 else "fp32"] = line
    ref = out["allow_tf32"].get("tasks_per_s")
    out["ratio_e2e_over_reference_tf32"] = round(our_e2e / ref, 2) if ref else None
    ref32 = out["fp32"].get("tasks_per_s")
    out["ratio_e2e_over_reference_fp32"] = round(our_e2e / ref32, 2) if ref32 else None
    out["what"] = ("unmodified reference run_train_iter on this GPU: sepconv K=5 256x448, 8 tasks per step, 3 warm-up "
                   "+ 5 timed steps, wall clock between torch.cuda.synchronize()")
    return out
