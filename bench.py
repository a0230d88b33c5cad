#!/usr/bin/env python
"""bench.py -- headline benchmark of the micrograph-denoiser inference path (BASELINE.json).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the hot path (the atrous Xception denoiser forward) over one batch of
synthetic 512x512 crops.  At N=1 the workload is BASELINE.json configs[1]: batch 32 of 512x512
crops on one B200 in the tensor-core mode that meets the parity contract (FP16 operands, FP32
accumulation; `dtype`).  N>1 (launched with torch.distributed.run, one rank per GPU) shards independent
crops across GPUs -- no data-path collective, weak scaling.

One JSON line on stdout (rank 0).  `value` = crops/s with the inputs resident in HBM; `e2e` = the
same metric through the public API with pinned HOST buffers (H2D and D2H inside the timed region);
`roofline` describes the dominant kernel, measured live with CUDA events; `cpu_baseline` is the
oracle (PyTorch-CPU restatement of the reference TF graph -- TensorFlow cannot run here) timed on
this box's host cores on a bounded sample.  `configs` carries the other BASELINE.json configs:
ms per 2048x2048 micrograph (GPU host-to-host / device-resident / in a stream, and the CPU port's 25
sequential batch-1 passes + stitch, DEN:666-677), the 96x96 x 4096 small-image path, batch-1 latency,
the other arithmetic modes; `stream_4096` is configs[3]: a stream of 4096x4096 micrographs, image k on
rank k % N, with a hash check that the sharded outputs equal a single GPU's bit for bit.

--impl reference times that CPU restatement alone, with all host threads, batch 1 per pass like
the reference's sess.run loop (DEN:646-647, 666-675).
"""
from __future__ import annotations

import argparse
import hashlib
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


def synthetic_micrograph(size, seed=1234):
    """One low-dose micrograph of size x size as raw Poisson counts (float32, NOT normalised: the wrapper's preprocess
    does that, DEN:655-656): smooth positive field x dose (gen_lq / get_scale, DMG:785-799)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:size, 0:size].astype(np.float32) / size
    f = np.full((size, size), 0.1, np.float32)
    for _ in range(24):
        cy, cx, sg, a = rng.random(), rng.random(), 0.02 + 0.2 * rng.random(), rng.random()
        f += a * np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * sg * sg))
    lam = 25.0 + rng.exponential(75.0)
    return rng.poisson(f / f.mean() * lam).astype(np.float32)


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


WORKLOAD = (f"batch {BATCH} of {CROP}x{CROP} synthetic crops per GPU per step through the atrous Xception denoiser "
            f"(variant A, random-init weights), BASELINE.json configs[1]")


def config_dict(n_sets):
    """Identical for both arms (the driver compares them): what is computed, not how."""
    return {"workload": WORKLOAD, "batch_per_gpu": BATCH, "crop": CROP,
            "l2": f"{n_sets} rotating input sets ({n_sets * BATCH * CROP * CROP * 4 / 1e6:.0f} MB > 126 MB L2); activations per step "
                  "(~1 GB per crop) far exceed L2; no explicit flush"}


def n_input_sets():
    return 5


def cpu_reference_leg(n_crops, s=CROP, stitch_2048=False):
    """The reference's CPU path through its PyTorch restatement (oracle/), all host threads, batch 1 per pass like the
    reference's sess.run loop (DEN:646-647).  Returns (crops/s from the median pass, threads, per-pass seconds, and -- with
    stitch_2048 -- ms for one 2048x2048 micrograph = 25 sequential passes + normalise / tile / stitch, DEN:653-682)."""
    import numpy as np
    import torch
    from oracle import wrapper as W
    from oracle.net import OracleNet
    emd = importlib.import_module("ai-cv-automation-elect-micr_b200")
    torch.set_num_threads(os.cpu_count() or 1)
    net = OracleNet(emd.weights.init_reference_weights(0), s)
    net.forward(synthetic_crops(1, s))          # warm-up
    ts = []
    ms_2048 = None
    if stitch_2048:
        img = synthetic_micrograph(2048, seed=5)

        def one(c):
            t0 = time.perf_counter()
            r = net.forward(c)
            ts.append(time.perf_counter() - t0)
            return r
        t0 = time.perf_counter()
        W.denoise(img, lambda crops: np.concatenate([one(crops[i:i + 1]) for i in range(len(crops))]), overlap=80, crop=s)
        ms_2048 = (time.perf_counter() - t0) * 1e3
    else:
        x = synthetic_crops(1, s)
        for _ in range(n_crops):
            t0 = time.perf_counter()
            net.forward(x)
            ts.append(time.perf_counter() - t0)
    med = sorted(ts)[len(ts) // 2]
    return 1.0 / med, torch.get_num_threads(), ts, ms_2048


def run_reference(args, rank, world):
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
    sample = (f"{done} batch-1 passes over one {CROP}x{CROP} crop of the workload's batch per timed step: the reference runs one "
              f"sess.run per crop (DEN:646-647), so its crops/s does not depend on the batch; PyTorch-CPU restatement of the "
              f"reference TF graph (TensorFlow is not installable here), FP32")
    emit({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / done, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(n_input_sets()),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def sha(arrays):
    h = hashlib.blake2b(digest_size=16)
    for a in arrays:
        h.update(a.tobytes() if hasattr(a, "tobytes") else a.numpy().tobytes())
    return h.hexdigest()


_REAL_STDOUT = None


def quiet_stdout():
    """The contract is ONE JSON line on stdout.  Native libraries write there too (NCCL prints its version banner from C on the
    first collective), so file descriptor 1 points at stderr for the whole run and the line goes to a duplicate of the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="fp16", choices=["bf16", "fp16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the other BASELINE configs (2048^2, 96^2, stream, modes)")
    ap.add_argument("--stream-images", type=int, default=8, help="4096x4096 micrographs per rank in the stream config")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
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
    S, B = CROP, BATCH
    eng = emd.Engine(device=local, cropsize=S, max_batch=B)
    eng.set_option("strict", 1)     # a GEMM-class layer with no tensor-core kernel is an error here, never a silent CUDA-core launch
    blob = emd.weights.pack(emd.weights.init_reference_weights(0))
    eng.load_weights(blob)
    tstream = torch.cuda.Stream()  # the engine launches on this stream; the timing events are recorded on it
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream

    # rotating input sets so the inputs alone exceed the 126 MB L2 (the activations, ~1 GB per crop, do anyway)
    n_sets = n_input_sets()
    host_sets = [torch.from_numpy(synthetic_crops(B, S, seed=1234 + 97 * rank + i)).pin_memory() for i in range(n_sets)]
    dev_sets = [h.cuda() for h in host_sets]
    d_out = torch.empty((B, S, S), dtype=torch.float32, device="cuda")
    h_out = torch.empty((B, S, S), dtype=torch.float32).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, finish=None):
        """W untimed calls, then exactly `steps` calls between CUDA events on the launching stream, barrier + synchronize on
        both sides, max over ranks.  `finish` (asynchronous calls): waits for everything outstanding, inside the timed region.
        Returns (ms total, kernel launches in the timed region summed over ranks)."""
        for i in range(warmup):
            fn(i)
        if finish:
            finish()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = eng.kernel_launches
        e0.record()
        for i in range(steps):
            fn(i)
        if finish:
            finish()
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

    dev_pass = lambda i: eng.forward(dev_sets[i % n_sets], out=d_out, mode=args.mode, stream=stream)
    host_pass = lambda i: eng.forward(host_sets[i % n_sets], out=h_out, mode=args.mode, stream=stream)
    h_outs = [h_out, torch.empty((B, S, S), dtype=torch.float32).pin_memory()]
    host_pass_async = lambda i: eng.forward(host_sets[i % n_sets], out=h_outs[i & 1], mode=args.mode, stream=stream, sync=False)

    sampler = ClockSampler(local)
    sampler.start()
    # pre-heat: ~2 s of the same passes before the W warm-up steps, so the K timed steps (0.2 s) run at the clocks a long job
    # settles at under the 1 kW cap, not at the boost clocks of an idle GPU
    t_heat = time.perf_counter()
    heat = 0
    while time.perf_counter() - t_heat < 2.0:
        for i in range(8):
            dev_pass(heat + i)
        torch.cuda.synchronize()
        heat += 8
    c0 = {k: eng.counter(k) for k in ("conv_cuda_core", "conv_fused_pair", "conv_fused_taps", "conv_fused_dw", "conv_tcgen05_gen1",
                                      "final_tcgen05", "tensor_core_launches", "launches")}
    ms, launches = timed(dev_pass, args.steps, args.warmup)
    per_pass = {k: (eng.counter(k) - c0[k]) / (args.steps + args.warmup) for k in c0}
    value = world * B * args.steps / (ms * 1e-3)
    if args.mode != "fp32":     # the benchmarked step ran on the kernels it claims: no CUDA-core conv, no first-generation kernel
        assert per_pass["conv_cuda_core"] == 0 and per_pass["conv_tcgen05_gen1"] == 0, per_pass
        assert per_pass["conv_fused_pair"] >= 39 and per_pass["conv_fused_dw"] >= 1 and per_pass["final_tcgen05"] == 1, per_pass

    # end to end through the public API with pinned host buffers: every step uploads its own batch and downloads its own result
    # inside the timed region.  A stream of batches is submitted the way a serving loop would: forward(sync=False) per batch
    # (emd_forward_async: batch i+1's upload and batch i-1's download run under batch i's pass) and one synchronize() at the end,
    # inside the timed region; `sync_value` is the same loop with a blocking call per batch.
    ms_e2e, _ = timed(host_pass_async, args.steps, 2, finish=lambda: eng.synchronize(stream))
    ms_e2e_sync, _ = timed(host_pass, args.steps, 2)
    clocks = sampler.stop()
    e2e = {"value": world * B * args.steps / (ms_e2e * 1e-3), "unit": UNIT,
           "h2d_bytes_per_step": B * S * S * 4, "d2h_bytes_per_step": B * S * S * 4,
           "api": "Engine.forward(pinned host in, pinned host out, sync=False) per step + Engine.synchronize() (emd_forward_async / emd_synchronize)",
           "sync_value": world * B * args.steps / (ms_e2e_sync * 1e-3)}

    out = {}
    pk = peaks()
    if rank == 0:
        # per-kernel device times (CUDA events around every step of the schedule, separate passes, mean of 3)
        eng.set_profile(True)
        acc = None
        for r in range(3):
            eng.forward(dev_sets[r % n_sets], out=d_out, mode=args.mode, stream=stream)
            cur = eng.step_info()
            acc = cur if acc is None else [(a[0], a[1] + c[1], a[2], a[3], a[4]) for a, c in zip(acc, cur)]
        info = [(n_, m_ / 3, f_, b_, l_) for n_, m_, f_, b_, l_ in acc]
        eng.set_profile(False)
        tot_ms = sum(i[1] for i in info)
        # dominant kernel = the step with the largest device time (one launch per step on this path)
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
        roof["survey_ideal_fusion_frac"] = (B / 4780.0 * 1e3) / (ms / args.steps)   # 4780 crops/s per B200 with every depthwise fused
        top = sorted(info, key=lambda i: -i[1])[:8]
        roof["top_steps_ms"] = {i[0]: round(i[1], 3) for i in top}
        out["roofline"] = roof
        out["kernels_per_pass"] = per_pass

    configs = {}
    stream_cfg = None
    if not args.no_configs:
        # ---- BASELINE configs[3]: a stream of 4096x4096 micrographs, image k -> rank k % N (sharding.denoise_stream's rule) ----
        M = args.stream_images
        bases = [torch.from_numpy(synthetic_micrograph(4096, seed=40 + j)).pin_memory() for j in range(4)]
        total = M * world
        mine = emd.sharding.shard_indices(total, rank, world)
        imgs = [bases[k % len(bases)] for k in mine]
        outs = [torch.empty((4096, 4096), dtype=torch.float32).pin_memory() for _ in range(M)]
        seng = emd.Engine(device=local, cropsize=S, max_batch=25)      # 100 crops per image = 4 balanced passes of 25
        seng.set_option("strict", 1)
        seng.load_weights(blob)
        run_stream = lambda ims: seng.denoise_images(ims, overlap=80, mode=args.mode, outs=outs[:len(ims)], out_dtype=np.float32, stream=stream)
        hashes = [sha([o]) for o in run_stream(imgs)]                  # warm-up pass (graph capture, buffers) + its hashes
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run_stream(imgs)
        e1.record()
        torch.cuda.synchronize()
        ms_stream = e0.elapsed_time(e1)
        hashes2 = [sha([o]) for o in outs[:len(imgs)]]
        if world > 1:
            t = torch.tensor([ms_stream], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_stream = float(t.item())
        # bit-identity: a second pass gives the same bits, and (N > 1) rank 0 recomputes every other rank's images alone on its
        # own GPU and compares the hashes -- the N-GPU result is the single-GPU result
        identical = hashes == hashes2
        if world > 1:
            gathered = [None] * world
            dist.all_gather_object(gathered, (mine, hashes2))
            if rank == 0:
                for r_mine, r_hashes in gathered[1:]:
                    res = run_stream([bases[k % len(bases)] for k in r_mine])
                    identical = identical and [sha([o]) for o in res] == r_hashes
        stream_cfg = {"workload": f"{M} micrographs of 4096x4096 per GPU, 100 overlapping 512x512 crops each (overlap 80), image k on rank k % N; "
                                  "pinned float32 in, float32 out",
                      "images": total, "images_per_s": total / (ms_stream * 1e-3), "crops_per_s": total * 100 / (ms_stream * 1e-3),
                      "ms_per_image_per_gpu": ms_stream / M,
                      "bit_identical_to_single_gpu": bool(identical) if rank == 0 else None}
        del seng

    if rank == 0 and not args.no_configs and world == 1:      # single-GPU configs: reported at N=1 only (the other ranks would idle)
        # ---- configs[2]: one 2048x2048 micrograph, 25 crops ----
        meng = emd.Engine(device=local, cropsize=S, max_batch=25)
        meng.set_option("strict", 1)
        meng.load_weights(blob)
        img = synthetic_micrograph(2048, seed=5)
        himg = torch.from_numpy(img).pin_memory()
        hout64 = torch.empty((2048, 2048), dtype=torch.float64).pin_memory()
        hout32 = torch.empty((2048, 2048), dtype=torch.float32).pin_memory()
        dimg = himg.cuda()
        dres = torch.empty((2048, 2048), dtype=torch.float64, device="cuda")

        def wall(fn, reps=10, warm=3, heat_s=1.0):
            """ms per call, host wall clock around `reps` calls + a final synchronize, after `warm` calls and `heat_s` seconds of
            the same calls (every figure of this section is taken at the clocks the GPU settles at under this load, like `value`)."""
            for _ in range(warm):
                fn()
            t_h = time.perf_counter()
            while time.perf_counter() - t_h < heat_s:
                fn()
                torch.cuda.synchronize()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            torch.cuda.synchronize()
            return (time.perf_counter() - t0) / reps * 1e3
        m = {"crops": 25, "overlap": 80,
             "gpu_e2e": wall(lambda: meng.denoise_image(himg, out=hout64, mode=args.mode)),
             "gpu_e2e_f32_out": wall(lambda: meng.denoise_image(himg, out=hout32, mode=args.mode, out_dtype=np.float32)),
             "gpu_device": wall(lambda: meng.denoise_image(dimg, out=dres, mode=args.mode)),
             "note": "ms, host wall clock around synchronous calls: gpu_e2e = pinned float32 image in, float64 image out on the host "
                     "(the reference's types, DEN:658); gpu_device = image and result resident in HBM; gpu_stream = per image in a stream "
                     "of 8 (copies of neighbouring images under the network passes, emd_denoise_stream)"}
        s_in = [himg] * 8
        s_out = [torch.empty((2048, 2048), dtype=torch.float32).pin_memory() for _ in range(8)]
        m["gpu_stream"] = wall(lambda: meng.denoise_images(s_in, outs=s_out, mode=args.mode, out_dtype=np.float32), reps=3, warm=1) / 8
        configs["ms_per_2048_micrograph"] = m
        # configs[0]-like latency: one crop, batch 1
        d1 = dev_sets[0][:1].contiguous()
        o1 = torch.empty_like(d1)
        configs["ms_per_512_crop_batch1"] = wall(lambda: meng.forward(d1, out=o1, mode=args.mode), reps=20, warm=5)
        del meng
        # the other arithmetic modes on configs[1]
        modes = {}
        for mode, steps in (("bf16", 20), ("fp16", 20), ("fp32", 2)):
            # (not timed(): that one is collective; this section runs on one rank only.)  All three modes the same way, so they
            # compare with each other; the headline `value` is the contract mode under the contract's timing rules.
            t_ms = wall(lambda mode=mode: eng.forward(dev_sets[0], out=d_out, mode=mode, stream=stream), reps=steps, warm=2,
                        heat_s=1.0 if mode != "fp32" else 0.0)
            modes[mode] = B / (t_ms * 1e-3)
        configs["crops_per_s_by_mode"] = modes
        # ---- configs[4]: 96x96 crops at batch 4096 (small_scans shape) ----
        del eng
        torch.cuda.empty_cache()
        e96 = emd.Engine(device=local, cropsize=96, max_batch=4096)
        e96.load_weights(blob)
        x96 = torch.from_numpy(np.random.default_rng(96).random((4096, 96, 96)).astype(np.float32)).cuda()
        y96 = torch.empty_like(x96)
        ms96 = wall(lambda: e96.forward(x96, out=y96, mode=args.mode, stream=stream), reps=3, warm=2)
        l0 = e96.kernel_launches
        e96.forward(x96, out=y96, mode=args.mode, stream=stream)
        torch.cuda.synchronize()
        configs["crops_96x96_batch4096"] = {"crops_per_s": 4096 / ms96 * 1e3, "ms_per_pass": ms96, "kernel_launches_per_pass": e96.kernel_launches - l0,
                                            "roofline_crops_per_s": 143000, "frac": 4096 / ms96 * 1e3 / 143000}
        del e96
        if not args.no_cpu_baseline and world == 1:
            # the CPU leg: ONE 2048x2048 micrograph through the oracle port = 25 sequential batch-1 passes + stitch (DEN:666-677);
            # its median pass also gives the CPU crops/s
            v, cores, ts, ms2048 = cpu_reference_leg(25, S, stitch_2048=True)
            configs["ms_per_2048_micrograph"]["cpu_port"] = ms2048
            configs["ms_per_2048_micrograph"]["cpu_cores"] = cores
            out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                   "sample": f"median of the 25 batch-1 passes of the PyTorch-CPU oracle over the 512x512 crops of one "
                                             f"2048x2048 micrograph (1 warm-up pass; the whole micrograph incl. stitch took {ms2048:.0f} ms)"}
    elif rank == 0 and not args.no_cpu_baseline and world == 1:
        v, cores, ts, _ = cpu_reference_leg(12, S)
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": f"median of 12 batch-1 passes of the PyTorch-CPU oracle over one {S}x{S} crop (1 warm-up)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.mode, "data": "synthetic", "config": config_dict(n_sets),
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "pre_heat": f"{heat} untimed passes (~2 s) before the warm-up steps: the timed steps run at steady-state clocks",
        }
        line.update(out)
        if configs:
            line["configs"] = configs
        if stream_cfg:
            line["stream_4096"] = stream_cfg
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
