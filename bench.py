#!/usr/bin/env python
"""Headline benchmark: fused SQ implicit-loss forward+backward throughput (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one pass of the hot path over one batch: ``ImplicitLoss(64, dev, 1.5, 260)(images, pred)`` followed by
``.backward()`` on BASELINE config 2 (batch 256, 64^3 grid, fp32) -- 67 108 864 points per GPU.  Under torchrun
(N > 1) every rank runs its own batch (the path shards by sample, no data-path collective: weak scaling) and the
step time is the max over ranks.

Printed line (rank 0), see the driver contract: ``value`` = Gpoints/s with inputs resident in HBM (CUDA events over
the K steps); ``e2e`` = the same through the host-buffer C-ABI call with H2D/D2H inside the timed region;
``roofline`` = the dominant kernel against the measured MUFU (SFU) peak of this GPU; ``cpu_baseline`` = the oracle
port on the box's host cores on a bounded sample.  ``--impl reference`` times that CPU path alone.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fused SQ implicit-loss fwd+bwd throughput"
UNIT = "Gpoints/s"
B, R, H, W = 256, 64, 256, 256           # BASELINE config 2; depth maps are 256x256 like the reference's dataset
TAU, SHARP = 1.5, 260.0                  # torch/train.py:64
MUFU_PER_POINT = 16                      # 12 (five pows + sigmoid) + 1 (exp(-tau cs)) + 3 (backward rcp)
CPU_SAMPLE_B = 8                         # bounded CPU sample: a B=8 slice of the same workload


def workload_config(n_gpus):
    return {"workload": f"ImplicitLoss(render_size={R}, tau={TAU}, sigmoid_sharpness={SHARP:g}) fwd+bwd, batch {B} per GPU, "
                        f"{R}^3 grid, depth maps {H}x{W} (BASELINE config 2)",
            "points_per_step_per_gpu": B * R ** 3, "batch_per_gpu": B, "render_size": R,
            "l2": "inputs rotate over 4 independent batches (4 x 67 MB of depth maps > 126 MB L2)",
            "reference_arm_sample": REFERENCE_SAMPLE,
            "parallelism": f"{n_gpus} x independent batch shards, no data-path collective"}


# ----------------------------------------------------------------------------------------------- CPU reference arm
REFERENCE_SAMPLE = (f"the reference arm and cpu_baseline run a B={CPU_SAMPLE_B} slice of this workload per step on the host cores "
                    f"({CPU_SAMPLE_B * R ** 3} points), scaled per point")


def cpu_reference(steps, warmup):
    """The reference's own ImplicitLoss (UNMODIFIED torch/classes.py staged under oracle/_ref by oracle/build_ref.py;
    the oracle port's loop form when that is absent) on all host cores, fwd+bwd, on a B=8 slice of the workload.
    This is the one place bench.py executes anything under oracle/ (the checker timed as the CPU baseline)."""
    from oracle import ref_import
    from oracle import sq_oracle as O
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    cpu = torch.device("cpu")
    true = O.random_params(CPU_SAMPLE_B, 0)
    pred = O.perturbed_params(true, 5)
    if ref_import.available() or ref_import.staged():
        rc, _ = ref_import.load()
        crit, kind = rc.ImplicitLoss(R, cpu, TAU, SHARP), "reference"
        what = f"UNMODIFIED reference torch/classes.py ImplicitLoss({R}, cpu, {TAU}, {SHARP:g}) (oracle/_ref)"
    else:
        crit, kind = O.ImplicitLoss(R, "cpu", TAU, SHARP, form="loop"), "port"
        what = "oracle/sq_oracle.py loop form (the reference's per-sample op sequence; oracle/_ref not staged)"
    with torch.no_grad():   # depth maps: the reference's own render of the true parameters at R, blown up to H x W so that
        small = crit.depth_projection(true).float()          # the nearest resize inside the loss returns it unchanged
    img = small.repeat_interleave(H // R, dim=1).repeat_interleave(W // R, dim=2).unsqueeze(1).contiguous()
    times = []
    for i in range(warmup + steps):
        p = pred.clone().requires_grad_(True)
        t0 = time.perf_counter()
        loss = crit(img, p)
        loss.backward()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    pts = CPU_SAMPLE_B * R ** 3
    mean = float(np.mean(times))
    cpu_name = ""
    try:
        cpu_name = [l.split(":")[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
    except Exception:
        pass
    return {"value": pts / mean / 1e9, "unit": UNIT, "cores": threads, "kind": kind,
            "sample": f"B={CPU_SAMPLE_B} slice of the workload ({pts} points per step), {steps} timed steps after {warmup} warm-up: "
                      f"mean {mean * 1e3:.1f} ms, best {np.min(times) * 1e3:.1f} ms per fwd+bwd; {what}; torch fp64, "
                      f"{threads} threads on {os.cpu_count()} host cores ({cpu_name})"}, mean


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    cb, mean = cpu_reference(steps, warmup)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": mean * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args.gpus),
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- helpers
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        busy = [v for v in sm if v > 0.5 * (max(mx) if mx else 1)] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_mufu_peak():
    """thread-MUFU-ops per clock per SM and the clock it was measured at, from the in-tree microbenchmark."""
    exe = os.path.join(ROOT, "sq_recovery_b200", "sq_peaks")
    try:
        out = subprocess.run([exe, "/dev/null", "--quick"], capture_output=True, text=True, timeout=120).stdout
        d = json.loads(out)
        return {"mufu_per_clk_sm": d["MUFU_MIX"]["mufu_per_clk_sm"], "sm_mhz": d["MUFU_MIX"]["sm_mhz"], "sms": d["sms"],
                "fp32_per_clk_sm": d["FFMA"]["fp32_per_clk_sm"], "source": "sq_peaks --quick, this run"}
    except Exception as e:   # fall back to the committed round-1 measurement
        try:
            d = json.load(open(os.path.join(ROOT, "profiles", "peaks_r01.json")))
            return {"mufu_per_clk_sm": d["MUFU_MIX"]["mufu_per_clk_sm"], "sm_mhz": d["MUFU_MIX"]["sm_mhz"], "sms": d["sms"],
                    "fp32_per_clk_sm": d["FFMA"]["fp32_per_clk_sm"], "source": f"profiles/peaks_r01.json ({type(e).__name__} running sq_peaks)"}
        except Exception:
            return {"mufu_per_clk_sm": 16.0, "sm_mhz": 1965.0, "sms": 148, "fp32_per_clk_sm": 128.0, "source": "nominal"}


def ncu_summary():
    """Issued XU (MUFU pipe) warp instructions and DRAM bytes per launch of the dominant kernel, from the committed ncu
    capture of this same seeded workload (tools/profile_step.py + tools/ncu_summary.py), or None."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "implicit_kernel_ncu_summary.json")))
    except Exception:
        return None


# ----------------------------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import torch.distributed as dist
    from sq_recovery_b200 import inputs as O                # seeded randsq / randquat workloads
    import sq_recovery_b200 as S
    from sq_recovery_b200 import _lib
    from sq_recovery_b200.functional import HostContext

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    steps, warmup = args.steps, max(args.warmup, 3)

    crit = S.ImplicitLoss(R, dev, TAU, SHARP)
    render = S.ImplicitLoss(H, dev, TAU, SHARP)             # synthetic depth maps = soft renders of the true params
    sets = []
    for k in range(4):                                      # 4 independent batches, rotated -> inputs exceed L2
        true = O.random_params(B, 1000 * rank + k)
        pred = O.perturbed_params(true, 7 + k).to(dev)
        img = render.depth_projection(true.to(dev)).unsqueeze(1).contiguous()
        sets.append((img, pred))
    torch.cuda.synchronize()

    def step(i):
        img, pred = sets[i % 4]
        p = pred.detach().requires_grad_(True)
        loss = crit(img, p)
        loss.backward()
        return loss, p.grad

    # The same step captured once per input set in a CUDA graph (public API inside the capture: the class call and
    # .backward()); replaying it removes the ~100 us of Python/autograd launch overhead per step, which is longer
    # than the kernels themselves.
    graphs = []
    if not args.eager:
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for k in range(4):
                for _ in range(2):
                    step(k)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize()
        for k in range(4):
            gph = torch.cuda.CUDAGraph()
            img, pred = sets[k]
            p = pred.detach().requires_grad_(True)
            with torch.cuda.graph(gph, stream=side):
                loss_k = crit(img, p)
                loss_k.backward()
            graphs.append((gph, loss_k, p))
        # ... and the four steps (one per input set) back to back in ONE graph: a single launch then covers ~220 us of
        # GPU work, so the host's launch rate (10-20 us per graph launch on these boxes, more on a busy host) cannot
        # leave the GPU idle between steps.  Same kernels, same work per step.
        quad = torch.cuda.CUDAGraph()
        quad_out = []
        with torch.cuda.graph(quad, stream=side):
            for k in range(4):
                img, pred = sets[k]
                p = pred.detach().requires_grad_(True)
                loss_k = crit(img, p)
                loss_k.backward()
                quad_out.append((loss_k, p))
        eager_step = step

        def step(i):                                        # noqa: F811  (graph replay of the step above)
            gph, loss_k, p = graphs[i % 4]
            gph.replay()
            return loss_k, p.grad

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for i in range(warmup):
        step(i)
    if graphs:
        for _ in range(3):
            quad.replay()
    # The timed region may last only milliseconds, shorter than nvidia-smi's sampling period, so the same step is
    # also run untimed for ~0.7 s right before it with the sampler on: the clocks / throttle reasons reported are
    # those of this workload under sustained load, and the timed steps follow back to back.
    t_soak = time.perf_counter()
    while time.perf_counter() - t_soak < 0.7:
        for i in range(50):
            step(i)
        torch.cuda.synchronize()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    if graphs:
        for _ in range(steps // 4):                         # four steps per graph launch ...
            quad.replay()
        loss, grad = quad_out[3][0], quad_out[3][1].grad
        for i in range(steps - steps % 4, steps):           # ... and the remainder one by one
            loss, grad = step(i)
    else:
        for i in range(steps):
            loss, grad = step(i)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    # dominant kernel alone, CUDA events recorded around its launch on its stream (sq_profile_events); eager launches
    kms = []
    one = eager_step if graphs else step
    for i in range(min(steps, 20)):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); b.record()                              # materialise the cudaEvent_t handles
        torch.cuda.synchronize()
        _lib.lib().sq_profile_events(a.cuda_event, b.cuda_event)
        one(i)
        torch.cuda.synchronize()
        kms.append(a.elapsed_time(b))
    # eager (no graph) step time for comparison
    eager_ms = None
    if graphs:
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(min(steps, 20)):
            eager_step(i)
        e1.record(); torch.cuda.synchronize()
        eager_ms = e0.elapsed_time(e1) / min(steps, 20)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    assert torch.isfinite(loss).item() and torch.isfinite(grad).all().item()

    # ---- e2e: host buffers -> C-ABI host call -> host results, copies inside the timed region
    ctx = HostContext(local)
    h_sets = []
    for img, pred in sets[:2]:
        hi = torch.empty(img.shape, dtype=torch.float32).pin_memory(); hi.copy_(img)
        hp = torch.empty(pred.shape, dtype=torch.float32).pin_memory(); hp.copy_(pred)
        h_sets.append((hi.numpy(), hp.numpy()))
    for i in range(3):
        ctx.implicit_loss(h_sets[i % 2][1], h_sets[i % 2][0], R, TAU, SHARP)
    barrier()
    e2e_steps = min(steps, 20)
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        l_h, g_h = ctx.implicit_loss(h_sets[i % 2][1], h_sets[i % 2][0], R, TAU, SHARP)
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = te.item()
    with torch.no_grad():                                   # the host path and the torch path agree
        l_t = crit(*sets[(e2e_steps - 1) % 2]).item()
        assert abs(l_h - l_t) <= 1e-6 * abs(l_t) and np.isfinite(g_h).all(), (l_h, l_t)
    ctx.close()

    if rank == 0:
        pts = B * R ** 3
        value = world * pts * steps / (ms * 1e-3) / 1e9
        peak = measured_mufu_peak()
        clock = clocks["sm_mhz"] or peak["sm_mhz"]
        kernel_ms = float(np.mean(kms))
        peak_gops = peak["mufu_per_clk_sm"] * peak["sms"] * peak["sm_mhz"] * 1e6 / 1e9       # thread-MUFU ops/s, measured
        ncu = ncu_summary()
        # ISSUED MUFU-pipe thread-ops per launch (ncu count for this seeded workload; includes the f64<->f32
        # conversions, which share the pipe) over the kernel time measured live = true pipe utilisation.
        issued = ncu["xu_warp_inst_per_launch"] * 32 if ncu else None
        achieved = issued / (kernel_ms * 1e-3) / 1e9 if issued else None
        dense_equiv = pts * MUFU_PER_POINT / (kernel_ms * 1e-3) / 1e9
        # Dispatch model measured with csrc/peaks.cu on this GPU (profiles/peaks_r01.json, EX2_FFMA{4,6,8}): per
        # scheduler a MUFU warp-instruction costs ~6 issue cycles and any other one ~0.9, and the XU itself 8 per MUFU:
        # t >= max(8 Nm, 6 Nm + 0.9 No) / (4 schedulers x SMs x clock).  This is the bound the kernel actually runs into.
        dispatch = None
        if ncu and ncu.get("warp_inst_per_launch"):
            nm, no = ncu["xu_warp_inst_per_launch"], ncu["warp_inst_per_launch"] - ncu["xu_warp_inst_per_launch"]
            cyc = max(8.0 * nm, 6.0 * nm + 0.9 * no) / (4 * peak["sms"])
            t_us = cyc / clock
            dispatch = {"mufu_warp_inst": nm, "other_warp_inst": no, "bound_us": t_us, "kernel_us": kernel_ms * 1e3,
                        "frac": t_us / (kernel_ms * 1e3),
                        "model": "max(8 Nm, 6 Nm + 0.9 No) issue cycles per scheduler (profiles/peaks_r01.json EX2_FFMA*)"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(world), launch="CUDA graph replay, four steps (one per input set) per graph launch" if graphs else "eager",
                           eager_ms_per_step=eager_ms),
            "clocks": clocks,
            "e2e": {"value": world * pts * e2e_steps / e2e_s / 1e9, "unit": UNIT,
                    # the pinned depth maps are sampled in place over PCIe: only the 32-byte sectors holding sampled
                    # pixels cross the bus (ncu dram_bytes_read of the same access pattern: 16.9 MB at 256 -> 64)
                    "h2d_bytes_per_step": B * R * min(W * 4, R * 32) + B * 12 * 4 + 4 * R * 4,
                    "d2h_bytes_per_step": 8 + B * 12 * 4, "host_image_bytes": B * H * W * 4,
                    "ms_per_step": e2e_s / e2e_steps * 1e3,
                    "api": "sq_implicit_loss_host (include/sqloss.h) on pinned host buffers"},
            "gpu_launches": 3 * steps,
            "roofline": {"bound": "sfu", "kernel": "implicit_kernel<true>", "achieved": achieved, "peak": peak_gops,
                         "unit": "G MUFU-op/s", "frac": (achieved / peak_gops) if achieved else None,
                         "traffic": ncu["dram_bytes_per_launch"] if ncu else None,
                         "kernel_ms": kernel_ms,
                         "how": "achieved = MUFU-pipe thread-ops the kernel ISSUES per launch (profiles/"
                                "implicit_kernel_ncu_summary.json) / live CUDA-event kernel time (events around the launch: "
                                "includes ~5 us of launch latency); the kernel skips grid points whose occupancy is below "
                                "2^-40 (box + ellipsoid bounds) and evaluates F with 8 MUFU ops instead of 10, so issued ops "
                                "are far fewer than the reference algorithm's 16 per grid point",
                         "dispatch_bound": dispatch,
                         "reference_algorithm_equivalent": {"mufu_per_point": MUFU_PER_POINT, "achieved": dense_equiv,
                                                            "frac": dense_equiv / peak_gops},
                         "peak_source": f"{peak['source']}: {peak['mufu_per_clk_sm']:.2f} MUFU/clk/SM x {peak['sms']} SMs x "
                                        f"{peak['sm_mhz']:.0f} MHz (of measured)",
                         "sm_mhz_during_run": clock},
        }
        if world == 1:
            line["cpu_baseline"] = cpu_reference(3, 1)[0]
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--eager", action="store_true", help="time eager launches instead of CUDA-graph replay")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
