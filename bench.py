"""Headline benchmark: septuplet-tasks/sec of the SepConv MAML inner loop (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one ``run_train_iter`` over a meta-batch of synthetic 256x448 septuplets:
per task K=5 inner steps (2 support triplets each: forward, L1, backward, fused update),
one query forward/backward, then the outer optimizer step (and, for N>1, the NCCL
all-reduce of the flat meta-gradient).  Prints ONE JSON line (see README / DESIGN.md).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

H, W = 256, 448
K_INNER = 5
TASKS_PER_GPU = 8
METRIC = "septuplet-tasks/sec (K=5 inner steps, 256x448)"
WORKLOAD = "sepconv MAML K=5 inner steps, 256x448 synthetic Vimeo-septuplet, meta-batch=8 per GPU, LSLR-SGD, 1*L1"


def bench_config(world):
    return {"workload": WORKLOAD, "tasks_per_gpu": TASKS_PER_GPU, "global_batch": TASKS_PER_GPU * world,
            "l2": "256 MB flush write between steps; per-step working set >> 126 MB L2"}


def make_args(batch, cuda=True):
    return argparse.Namespace(
        model='sepconv', loss='1*L1', optimizer='SGD', inner_lr=1e-5, outer_lr=1e-5, batch_size=batch, mode='train',
        resume=True, number_of_training_steps_per_iter=K_INNER, number_of_evaluation_steps_per_iter=K_INNER,
        metasgd=False, attenuate=False, learnable_per_layer_per_step_inner_loop_learning_rate=False,
        enable_inner_loop_optimizable_bn_params=False, second_order=False, first_order_to_second_order_epoch=-1,
        use_multi_step_loss_optimization=False, multi_step_loss_num_epochs=1, random_seed=12345, cuda=cuda,
        num_gpu=1 if cuda else 0, pretrained_model=None, weight_decay=1e-4,
        load_checkpoint=False)   # seeded random init: no checkpoint files on the box


def synthetic_septuplets(batch, seed, h=H, w=W):
    """Smooth random texture translated by a per-task velocity + noise, clamp [0,1]."""
    g = torch.Generator().manual_seed(seed)
    frames = [torch.empty(batch, 3, h, w) for _ in range(7)]
    for b in range(batch):
        base = torch.rand(1, 3, h // 8 + 8, w // 8 + 8, generator=g)
        big = torch.nn.functional.interpolate(base, scale_factor=8, mode='bilinear', align_corners=False)[0]
        vx, vy = (torch.rand(2, generator=g) * 6 - 3).tolist()
        for t in range(7):
            ox, oy = 32 + int(round(vx * t)), 32 + int(round(vy * t))
            crop = big[:, oy:oy + h, ox:ox + w]
            frames[t][b] = (crop + 0.02 * torch.randn(3, h, w, generator=g)).clamp(0, 1)
    return frames


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.stop_flag, self.thread = index, [], False, None

    def _loop(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:
                pass
            time.sleep(0.2)

    def start(self):
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=6)
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.samples)}


def dist_setup(n_gpus):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local),
                                timeout=datetime.timedelta(seconds=180))
    else:
        torch.cuda.set_device(0)
    return rank, world, local


def run_ours(a):
    import torch.distributed as dist
    from meta_interpolation_b200.meta_learning_system import SceneAdaptiveInterpolation
    rank, world, local = dist_setup(a.gpus)
    batch = TASKS_PER_GPU * world
    system = SceneAdaptiveInterpolation(make_args(batch))
    ops = system.ops
    n_sets = 2
    host_sets = [[f.pin_memory() for f in synthetic_septuplets(batch, 100 + s)] for s in range(n_sets)]
    dev_sets = [[f.cuda(non_blocking=True) for f in hs] for hs in host_sets]
    l2_flush = torch.empty(64 * 1024 * 1024, device="cuda", dtype=torch.float32)   # 256 MB > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """K steps bracketed by barrier+sync, device time from CUDA events, max over ranks."""
        barrier()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for i in range(steps):
            fn(i)
        stop.record()
        barrier()
        ms = torch.tensor([start.elapsed_time(stop)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- device-resident leg (value)
    def step_resident(i):
        l2_flush.zero_()
        system.run_train_iter(dev_sets[i % n_sets], epoch=0)

    for i in range(a.warmup):
        step_resident(i)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = ops.launch_count()
    ms = timed(step_resident, a.steps)
    launches = ops.launch_count() - n0
    clocks = sampler.stop() if rank == 0 else None
    value = batch * a.steps / (ms / 1e3)

    # ---- end-to-end leg (pinned host frames -> device every step, loss read back every step)
    h2d = sum(f.numel() * 4 for f in host_sets[0])   # whole job: each rank copies the 1/world of it that it adapts
    loss_host = torch.zeros(1).pin_memory()
    per_rank = batch // world
    preds_host = torch.zeros(per_rank, 3, H, W).pin_memory()   # run_train_iter's second result: the rank's predictions

    def step_e2e(i):
        l2_flush.zero_()
        per = batch // world
        hs = host_sets[i % n_sets]
        frames = [torch.empty(batch, 3, H, W, device="cuda") for _ in range(7)]
        for t in range(7):
            frames[t][rank * per:(rank + 1) * per].copy_(hs[t][rank * per:(rank + 1) * per], non_blocking=True)
        losses, preds, _ = system.run_train_iter(frames, epoch=0)
        loss_host.copy_(losses['loss'].detach().reshape(1), non_blocking=True)
        for j in range(per_rank):          # the predictions of the tasks this rank adapted, back to the host
            preds_host[j].copy_(preds[rank * per_rank + j].reshape(3, H, W), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    step_e2e(0)
    ms_e2e = timed(step_e2e, a.steps)
    e2e = batch * a.steps / (ms_e2e / 1e3)

    # the instrumented roofline iteration contains the meta-gradient all-reduce: every rank must take part
    extra = extra_sections(system, a, rank)
    del system, dev_sets
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    other = None
    if not a.no_other_configs:
        from bench_sections import other_configs_section
        other = other_configs_section(rank, world, timed)
    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 4), "unit": "tasks/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": round(ms / a.steps, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 (tf32 tensor-core conv, fp32 accumulate)", "data": "synthetic",
            "config": bench_config(world),
            "clocks": clocks,
            "e2e": {"value": round(e2e, 4), "unit": "tasks/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": world * (4 + int(preds_host.numel() * 4)),
                    "bytes_scope": "all ranks together; result read back = loss + every prediction [3,256,448]"},
            "gpu_launches": int(launches),
        }
        line.update(extra)
        if other is not None:
            line["config"]["other_configs"] = other
        if world == 1 and not a.no_gpu_reference:
            from bench_sections import gpu_reference_section
            line["gpu_reference"] = gpu_reference_section(line["e2e"]["value"])
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def extra_sections(system, a, rank):
    out = {}
    try:
        from bench_sections import roofline_section, cpu_baseline_section
        out["roofline"] = roofline_section(system)
        if rank == 0 and int(os.environ.get("WORLD_SIZE", "1")) == 1 and not a.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline_section()
    except ImportError:
        pass
    return out


def run_reference(a):
    """The reference's CPU implementation of the path = the pinned oracle port, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from bench_sections import cpu_reference_run
    print(json.dumps(cpu_reference_run(a.steps, a.warmup, int(os.environ.get("WORLD_SIZE", str(a.gpus))))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", dest="no_cpu_baseline", action="store_true")
    ap.add_argument("--no-other-configs", dest="no_other_configs", action="store_true",
                    help="skip the C3/C4/C5 legs (BASELINE configs[2..4]) after the timed region")
    ap.add_argument("--no-gpu-reference", dest="no_gpu_reference", action="store_true",
                    help="skip timing the unmodified reference's own GPU path (baseline/_ref) at N=1")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
