#!/usr/bin/env python
"""bench.py -- headline benchmark of the micrograph-denoiser inference path (BASELINE.json).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the hot path (the atrous Xception denoiser forward) over one batch of
synthetic 512x512 crops.  At N=1 the workload is BASELINE.json configs[1]: batch 32 of 512x512
crops on one B200 in the BF16 tensor-core mode.  N>1 (launched with torch.distributed.run, one rank
per GPU) shards independent crops across GPUs -- no data-path collective, weak scaling.

One JSON line on stdout (rank 0).  `value` = crops/s with the inputs resident in HBM; `e2e` = the
same metric through the public API with pinned HOST buffers (H2D and D2H inside the timed region);
`roofline` describes the dominant kernel, measured live with CUDA events; `cpu_baseline` is the
oracle (PyTorch-CPU restatement of the reference TF graph -- TensorFlow cannot run here) timed on
this box's host cores on a bounded sample.

--impl reference times that CPU restatement alone, with all host threads, batch 1 per pass like
the reference's sess.run loop (DEN:646-647, 666-675).
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "denoised 512x512 crops/s"
UNIT = "crops/s"
CROP = 512
BATCH = 32


def synthetic_crops(n, s, seed=1234):
    """Poisson low-dose crops like the reference's generator (gen_lq, DMG:789-799; dose law DMG:785-786):
    smooth positive field x dose -> Poisson -> scale0to1."""
    import numpy as np
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:s, 0:s].astype(np.float32) / s
    out = np.empty((n, s, s), np.float32)
    for i in range(n):
        f = np.full((s, s), 0.1, np.float32)
        for _ in range(8):
            cy, cx, sg, a = rng.random(), rng.random(), 0.05 + 0.25 * rng.random(), rng.random()
            f += a * np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * sg * sg))
        lam = 25.0 + rng.exponential(75.0)
        lq = rng.poisson(f / f.mean() * lam).astype(np.float32)
        mn, mx = lq.min(), lq.max()
        out[i] = (lq - mn) / (mx - mn) if mx > mn else 0.5
    return out


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tflops": d["bf16_tflops_sustained"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def measured_traffic(layer, batch):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernel that runs `layer`, from the committed
    `ncu --set full` capture (profiles/traffic.json, written by tools/ncu_traffic.py), scaled to this batch."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None
    d = json.load(open(p)).get(layer)
    if not d:
        return None
    return d["dram_bytes_per_launch"] * batch / d["batch"]


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._halt = threading.Event()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            }
            while not self._halt.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.05)
        except Exception as ex:  # NVML missing: report that rather than fail the bench
            self.reasons.add(f"nvml_unavailable:{type(ex).__name__}")

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def time_cpu_oracle(n_pass, s=CROP, seed=0):
    """Bounded CPU sample: n_pass batch-1 passes of the oracle after one warm-up.  Returns crops/s."""
    import torch
    from oracle.net import OracleNet
    emd = importlib.import_module("ai-cv-automation-elect-micr_b200")
    torch.set_num_threads(os.cpu_count() or 1)
    net = OracleNet(emd.weights.init_reference_weights(seed), s)
    x = synthetic_crops(1, s)
    net.forward(x)
    ts = []
    for _ in range(n_pass):
        t0 = time.perf_counter()
        net.forward(x)
        ts.append(time.perf_counter() - t0)
    ts.sort()
    return 1.0 / ts[len(ts) // 2], torch.get_num_threads()


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path (oracle port), all host threads."""
    if rank != 0:
        return
    t0 = time.perf_counter()
    import torch
    from oracle.net import OracleNet
    emd = importlib.import_module("ai-cv-automation-elect-micr_b200")
    torch.set_num_threads(os.cpu_count() or 1)
    net = OracleNet(emd.weights.init_reference_weights(0), CROP)
    x = synthetic_crops(1, CROP)
    for _ in range(max(args.warmup, 1)):
        net.forward(x)
        if time.perf_counter() - t0 > 120:
            break
    steps = args.steps
    t1 = time.perf_counter()
    done = 0
    for _ in range(steps):
        net.forward(x)
        done += 1
        if time.perf_counter() - t1 > 150:  # keep the whole run within a few minutes
            break
    dt = time.perf_counter() - t1
    v = done / dt
    cores = torch.get_num_threads()
    sample = f"{done} batch-1 passes over one 512x512 crop (the reference runs one sess.run per crop, DEN:646-647)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / done, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"batch {BATCH} of {CROP}x{CROP} synthetic crops per GPU per step through the atrous Xception "
                               f"denoiser (variant A, random-init weights), BASELINE.json configs[1]",
                   "batch_per_gpu": BATCH, "crop": CROP, "mode": "f32",
                   "sample_per_step": "1 of the batch's crops per timed step (the reference runs one sess.run per crop, "
                                      "DEN:646-647, so crops/s does not depend on the batch); PyTorch-CPU restatement of the "
                                      "reference TF graph, TensorFlow is not installable here"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="bf16", choices=["bf16", "fp16", "fp32"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--crop", type=int, default=CROP)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    emd = importlib.import_module("ai-cv-automation-elect-micr_b200")
    args.warmup = max(args.warmup, 3)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    S, B = args.crop, args.batch
    eng = emd.Engine(device=local, cropsize=S, max_batch=B)
    eng.load_weights(emd.weights.pack(emd.weights.init_reference_weights(0)))
    tstream = torch.cuda.Stream()  # the engine launches on this stream; the timing events are recorded on it
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream

    # rotating input sets so the inputs alone exceed the 126 MB L2 (the activations, ~1 GB per crop, do anyway)
    n_sets = max(2, int(np.ceil(140e6 / (B * S * S * 4))))
    host_sets = [torch.from_numpy(synthetic_crops(B, S, seed=1234 + 97 * rank + i)).pin_memory() for i in range(min(n_sets, 5))]
    dev_sets = [h.cuda() for h in host_sets]
    d_out = torch.empty((B, S, S), dtype=torch.float32, device="cuda")
    h_out = torch.empty((B, S, S), dtype=torch.float32).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = eng.kernel_launches
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        launches = eng.kernel_launches - l0
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            lt = torch.tensor([launches], device="cuda", dtype=torch.int64)
            dist.all_reduce(lt)
            launches = int(lt.item())
        barrier()
        return ms, launches

    sampler = ClockSampler(local)
    sampler.start()
    ms, launches = timed(lambda i: eng.forward(dev_sets[i % len(dev_sets)], out=d_out, mode=args.mode, stream=stream),
                         args.steps, args.warmup)
    clocks = sampler.stop()
    value = world * B * args.steps / (ms * 1e-3)

    # end to end through the public API with pinned host buffers (H2D + D2H inside the timed region)
    ms_e2e, _ = timed(lambda i: eng.forward(host_sets[i % len(host_sets)], out=h_out, mode=args.mode, stream=stream),
                      args.steps, 1)
    e2e = {"value": world * B * args.steps / (ms_e2e * 1e-3), "unit": UNIT,
           "h2d_bytes_per_step": B * S * S * 4, "d2h_bytes_per_step": B * S * S * 4}

    # per-kernel device times (CUDA events around every step of the schedule, separate pass)
    out = {}
    if rank == 0:
        pk = peaks()
        eng.set_profile(True)
        eng.forward(dev_sets[0], out=d_out, mode=args.mode, stream=stream)
        info = eng.step_info()
        eng.set_profile(False)
        tot_ms = sum(i[1] for i in info)
        # dominant kernel = the step with the largest device time; a step is one kernel launch, except the
        # transposed convs (4 sub-pixel phase launches of the same kernel): per-launch figures divide by `launches`
        name, kms, fl, by, nl = max(info, key=lambda i: i[1])
        nl = max(nl, 1)
        fl, by = fl * B / nl, by * B / nl          # algorithmic work of ONE launch (whole batch)
        lms = kms / nl                              # average launch duration (CUDA events on the launching stream)
        t_tensor, t_hbm = fl / (pk["tflops"] * 1e12), by / (pk["hbm_gbs"] * 1e9)
        if t_tensor >= t_hbm:
            roof = {"bound": "tensor", "achieved": fl / (lms * 1e-3) / 1e12, "peak": pk["tflops"], "unit": "TFLOP/s"}
        else:
            roof = {"bound": "hbm", "achieved": by / (lms * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s"}
        roof["frac"] = roof["achieved"] / roof["peak"]
        roof["traffic"] = measured_traffic(name, B)
        roof["kernel"] = name
        roof["launches_per_step"] = nl
        roof["kernel_ms"] = lms
        roof["kernel_share_of_step"] = kms / tot_ms
        roof["peak_source"] = pk["source"] + " (sustained BF16 figure: the kernel is timed inside a long step)"
        # whole-network roofline with per-layer fusion: sum over steps of max(FLOPs/P_tensor, bytes/BW_hbm)
        # with the ALGORITHMIC bytes of the steps as they ran (depthwise computed inside a GEMM kernel: its
        # intermediate is not counted)
        t_roof = sum(max(f * B / (pk["tflops"] * 1e12), b * B / (pk["hbm_gbs"] * 1e9)) for _, _, f, b, n_ in info if n_ > 0)
        roof["network_roofline_ms"] = t_roof * 1e3
        roof["network_frac"] = t_roof * 1e3 / (ms / args.steps)
        top = sorted(info, key=lambda i: -i[1])[:8]
        roof["top_steps_ms"] = {i[0]: round(i[1], 3) for i in top}
        out["roofline"] = roof
        if not args.no_cpu_baseline and world == 1:   # the CPU leg is reported at N=1 only (the other ranks would wait for it)
            n_pass = 12   # ~10 s of CPU work on the box's host cores
            v, cores = time_cpu_oracle(n_pass, S)
            out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                   "sample": f"median of {n_pass} batch-1 passes of the PyTorch-CPU oracle over one "
                                             f"{S}x{S} crop (1 warm-up)"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.mode, "data": "synthetic",
            "config": {"workload": f"batch {B} of {S}x{S} synthetic crops per GPU per step through the atrous Xception "
                                   f"denoiser (variant A, random-init weights), BASELINE.json configs[1]",
                       "batch_per_gpu": B, "crop": S, "mode": args.mode,
                       "l2": f"{len(dev_sets)} rotating input sets ({len(dev_sets) * B * S * S * 4 / 1e6:.0f} MB > 126 MB L2); "
                             "activations per step far exceed L2; no explicit flush"},
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        }
        line.update(out)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
