#!/usr/bin/env python
"""Headline benchmark: SuperResolutionNet x2 training throughput (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle port)

Workload (N=1 and per GPU for N>1, weak scaling): SuperResolutionNet(scale=2, 64 features, 8 dense
blocks, temporal_window=1), synthetic 640x360 -> 1280x720 clips, T=3, B=16 per GPU, bf16 activations,
one step = forward + MSE loss + backward (+ bucketed gradient all-reduce) + AdamW.

Prints ONE JSON line on rank 0.  `value`: clips ("frames" = output HR centre frames) per second with
inputs resident in HBM; `e2e`: the same through the public module API with the step's inputs copied
from pinned host memory and the loss read back, inside the timed region; `roofline`: the dense-conv
kernel family, algorithmic FLOPs / CUDA-event time per launch summed over the timed steps, against the
measured cuBLAS bf16 peak in MEASURED_PEAKS.json; `cpu_baseline`: the oracle port on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "continual-learning-for-dynamic-video-quality-enhancement_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "sr_x2_train_frames_per_sec"
UNIT = "frames/s"
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["native", "reference"], default="native")
    # workload overrides (debugging only; the defaults ARE the benchmark configuration)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--height", type=int, default=360)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--features", type=int, default=64)
    ap.add_argument("--blocks", type=int, default=8)
    ap.add_argument("--dtype", choices=["bf16", "fp32"], default="bf16")
    ap.add_argument("--engine", choices=["auto", "simt", "tc"], default="auto")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-eager", action="store_true")
    ap.add_argument("--value-only", action="store_true",
                    help="warm-up + the K device-resident steps only (what the ncu passes of scripts/gpu_round.sh run)")
    ap.add_argument("--cpu-budget", type=float, default=25.0, help="seconds of CPU work for cpu_baseline")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        d["_source"] = "measured"
        return d
    d = dict(FALLBACK_PEAKS)
    d["_source"] = "fallback"
    return d


def synth_batch(B, T, H, W, scale, seed, device):
    """SURVEY.md section 8d synthetic clips: uniform centre frame, neighbours = centre rolled by an
    integer shift in [-3,3]^2 plus N(0, 0.01^2) noise, clamped to [0,1]; uniform HR target."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    centre = torch.rand(B, 3, H, W, generator=g, device=device)
    frames = []
    for t in range(T):
        if t == T // 2:
            frames.append(centre)
            continue
        dx, dy = (t * 5 + 1) % 7 - 3, (t * 3 + 2) % 7 - 3
        f = torch.roll(centre, (dy, dx), (2, 3)) + 0.01 * torch.randn(B, 3, H, W, generator=g, device=device)
        frames.append(f.clamp_(0, 1))
    lr = torch.stack(frames, 1).contiguous()
    hr = torch.rand(B, 3, H * scale, W * scale, generator=g, device=device)
    return lr, hr


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU every 200 ms through NVML while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        med = s[len(s) // 2] if s else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm on the host cores (oracle port; /root/reference is not on the box)
# ------------------------------------------------------------------------------------------------
def _reference_module():
    """The UNMODIFIED reference package (oracle/_ref, installed by oracle/build_ref.py) or None."""
    from oracle import build_ref
    path = build_ref.ref_path()
    if path is None:
        return None
    if path not in sys.path:
        sys.path.insert(0, path)
    import nerve_cl.models as ref_models
    return ref_models


class CpuArm:
    """One fwd + MSE + bwd step of the reference's SuperResolutionNet on the host cores.  Runs the reference's own
    nn.Module when oracle/_ref exists (kind "reference"), else the oracle port over a random state_dict of the
    same shapes (kind "port").  Never touches the product package or the GPU."""

    def __init__(self, features, blocks, scale=2, tw=1):
        import torch
        from oracle import sr_oracle
        self.scale, self.T = scale, 2 * tw + 1
        ref = _reference_module()
        torch.manual_seed(0)
        if ref is not None:
            self.kind = "reference"
            self.model = ref.SuperResolutionNet(scale_factor=scale, num_features=features, num_residual_blocks=blocks,
                                                temporal_window=tw).train()
        else:
            self.kind = "port"
            self.sd = sr_oracle.random_state_dict(scale, features, blocks, self.T)

    def step_seconds(self, B, H, W, reps=1):
        import torch
        from oracle import sr_oracle
        lr, hr = synth_batch(B, self.T, H, W, self.scale, 1234, "cpu")
        best = float("inf")
        for _ in range(reps):
            t0 = time.perf_counter()
            if self.kind == "reference":
                self.model.zero_grad()
                torch.nn.functional.mse_loss(self.model(lr), hr).backward()
            else:
                sr_oracle.train_step_grads(self.sd, lr, hr, self.scale, True)
            best = min(best, time.perf_counter() - t0)
        return best


def pick_cpu_sample(arm, args, budget_s):
    """Choose how many rows of one 640-wide clip fit `budget_s` seconds of CPU work per step, from a probe."""
    probe_h = 32
    t = arm.step_seconds(1, probe_h, args.width)          # includes first-call warm-up
    t = min(t, arm.step_seconds(1, probe_h, args.width))
    rows = int(budget_s / max(t, 1e-6) * probe_h)
    rows = max(16, min(args.height, rows // 8 * 8))
    return rows


def lscpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_cfg1(arm, iters=5):
    """BASELINE.json configs[0]: B=4, 3x64x64 windows, fwd+bwd on the CPU: median and min of `iters` timed steps."""
    arm.step_seconds(4, 64, 64)
    ts = sorted(arm.step_seconds(4, 64, 64) for _ in range(iters))
    return {"shape": [4, 3, 3, 64, 64], "iters": iters, "median_s": ts[len(ts) // 2], "min_s": ts[0],
            "frames_per_s_median": 4.0 / ts[len(ts) // 2]}


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    arm = CpuArm(args.features, args.blocks)
    total = max(args.steps + args.warmup, 1)
    rows = pick_cpu_sample(arm, args, budget_s=max(120.0 / total, 2.0))
    frac = rows / args.height
    for _ in range(args.warmup):
        arm.step_seconds(1, rows, args.width)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        arm.step_seconds(1, rows, args.width)
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    value = frac / dt
    what = "the reference's own SuperResolutionNet (oracle/_ref)" if arm.kind == "reference" else "oracle port"
    sample = (f"{what}: 1 clip x {rows}/{args.height} rows of the {args.width}x{args.height} LR window per step "
              f"(T=3, fwd+MSE+bwd, fp32, torch CPU {torch.get_num_threads()} threads), scaled by pixel count; the rate "
              f"is per clip, so the arm's B=1 sample compares with the native arm's B={args.batch}")
    cfg = workload_config(args, args.gpus)
    cfg["reference_arm_batch"] = 1
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3 / frac, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": arm.kind, "sample": sample,
                         "cpu": lscpu_model(), "cfg1": cpu_cfg1(arm)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, n):
    return {
        "workload": (f"SuperResolutionNet x2 training, {args.width}x{args.height}->{2 * args.width}x{2 * args.height}, "
                     f"T=3, {args.features} feat, {args.blocks} dense blocks, B={args.batch}/GPU"),
        "global_batch": args.batch * n, "frames_per_window": 3, "parallelism": f"dp{n}",
        "l2": "working set (tens of GB of activations per step) >> 126 MB L2; no explicit flush needed",
        "optimizer": "AdamW (fused nervecl kernel over the flat parameter buffer)",
    }


def gpu_eager(dev, args, batch=2, iters=5):
    """Same-box GPU comparator: the UNMODIFIED reference module (oracle/_ref; the oracle port
    when it is absent) through ATen/cuDNN eager on this B200 -- fp32 with TF32 off and bf16 autocast -- on `batch`
    windows of the benchmark shape (eager autograd keeps ~6.4 GB of fp32 activations per 360p sample), one
    fwd + MSE + bwd per step, rate per clip.  None of our kernels run here."""
    import torch
    from oracle import sr_oracle
    ref = _reference_module()
    torch.manual_seed(0)
    if ref is not None:
        model = ref.SuperResolutionNet(scale_factor=2, num_features=args.features, num_residual_blocks=args.blocks,
                                       temporal_window=1).to(dev).train()
        kind = "reference"
    else:
        sd = {k: v.to(dev) for k, v in sr_oracle.random_state_dict(2, args.features, args.blocks, 3).items()}
        kind = "port"
    lr, hr = synth_batch(batch, 3, args.height, args.width, 2, 77, dev)
    out = {"impl": kind, "batch": batch, "iters": iters,
           "note": "ATen/cuDNN eager fwd+MSE+bwd of the reference module on this GPU; frames/s = clips/s"}
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        for tag, ac in (("fp32_tf32_off", False), ("bf16_autocast", True)):
            def one():
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=ac):
                    if kind == "reference":
                        model.zero_grad()
                        torch.nn.functional.mse_loss(model(lr).float(), hr).backward()
                    else:
                        sr_oracle.train_step_grads(sd, lr, hr, 2, True)
            try:
                one(); one()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(iters):
                    one()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / iters
                out[tag] = {"ms_per_step": ms, "frames_per_s": batch / (ms / 1e3)}
            except Exception as exc:       # (an eager OOM must not take the benchmark line down)
                out[tag] = {"error": f"{type(exc).__name__}: {str(exc)[:160]}"}
                torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    if kind == "reference":
        del model
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------
# native arm
# ------------------------------------------------------------------------------------------------
def run_native(args):
    import torch
    import torch.distributed as dist
    from nerve_cl_b200 import ops
    from nerve_cl_b200.engine import KernelTimer
    from nerve_cl_b200.models import SuperResolutionNet
    from nerve_cl_b200.optim import FlatAdamW
    from nerve_cl_b200 import distributed as nd

    rank, local_rank, world = nd.init_from_env()
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (native arm) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    B, T, H, W, s = args.batch, 3, args.height, args.width, 2

    torch.manual_seed(0)
    model = SuperResolutionNet(scale_factor=s, num_features=args.features, num_residual_blocks=args.blocks,
                               temporal_window=1).to(dev).train()
    model.compute_dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    model.conv_engine = {"auto": ops.CONV_AUTO, "simt": ops.CONV_SIMT, "tc": ops.CONV_TC}[args.engine]
    nd.data_parallel(model)
    opt = FlatAdamW(model, lr=1e-3, weight_decay=1e-5)

    lr_dev, hr_dev = synth_batch(B, T, H, W, s, 1234 + rank, dev)
    lr_host = lr_dev.cpu().pin_memory()
    hr_host = hr_dev.cpu().pin_memory()
    h2d = lr_host.numel() * 4 + hr_host.numel() * 4

    def step(lr, hr):
        opt.zero_grad()
        out = model(lr)
        loss = torch.nn.functional.mse_loss(out, hr)
        loss.backward()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    timing_notes = []

    def timed(fn, steps):
        """K calls of fn between two CUDA events on the launching stream, barrier + synchronize on both sides, max over
        ranks.  The event time is cross-checked against the host clock around the same region (which brackets it from
        above): a pass whose two clocks disagree by more than 5 % is repeated (up to 3 times) and noted."""
        for attempt in range(3):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            barrier()
            wall = (time.perf_counter() - t0) * 1e3
            dev_ms = e0.elapsed_time(e1)
            bad = torch.tensor([0.0 if abs(wall - dev_ms) <= 0.05 * wall + 2.0 else 1.0], device=dev)
            if world > 1:
                dist.all_reduce(bad, op=dist.ReduceOp.MAX)       # every rank takes the same decision (collectives must pair)
            if float(bad.item()) == 0.0:
                break
            timing_notes.append({"attempt": attempt, "event_ms": dev_ms, "wall_ms": wall})
        ms = torch.tensor([dev_ms], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(max(args.warmup, 3)):
        step(lr_dev, hr_dev)

    # ---- device-resident throughput (the line's `value`): K steps, nothing but the step in the timed region ----
    launches0 = ops.LAUNCHES[0]
    sampler = ClockSampler(local_rank)
    sampler.start()
    # three back-to-back passes of K steps; `value` is their MEDIAN, all three are reported
    value_passes = [timed(lambda: step(lr_dev, hr_dev), args.steps)]
    launches = ops.LAUNCHES[0] - launches0
    if not args.value_only:                      # (the ncu passes run --value-only: one pass keeps them short)
        value_passes += [timed(lambda: step(lr_dev, hr_dev), args.steps) for _ in range(2)]
    ms = sorted(value_passes)[len(value_passes) // 2]
    clocks = sampler.stop()

    if args.value_only:
        if rank == 0:
            print(json.dumps({"value_only": True, "ms_per_step": ms / args.steps, "steps": args.steps,
                              "gpu_launches": launches}))
        return

    # ---- the same K steps again with per-launch CUDA-event spans (kernel shares, roofline numerators) ----
    plan = next(iter(model._plans.values()))
    timer = KernelTimer()
    plan.timer = timer
    import nerve_cl_b200.engine as _eng
    _eng.nv.timer = timer
    ms_spans = timed(lambda: step(lr_dev, hr_dev), args.steps)
    plan.timer = None
    _eng.nv.timer = None
    kdetail = timer.summary()
    ksum = {}
    for k, d in kdetail.items():
        agg = ksum.setdefault(k.split("|")[0], {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
        for f in agg:
            agg[f] += d[f]
    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "bench_detail.json"), "w") as f:
            json.dump({k: {"launches_per_step": d["launches"] / args.steps, "ms_per_step": d["ms"] / args.steps,
                           "tflops": d["flops"] / (d["ms"] / 1e3) / 1e12 if d["ms"] > 0 and d["flops"] else None}
                       for k, d in sorted(kdetail.items(), key=lambda kv: -kv[1]["ms"])}, f, indent=1)

    # ---- end to end: pinned host -> device copies and loss read-back inside the timed region ----
    losses = []

    # The loader a user writes: pinned host batches, non_blocking copies on a side stream, the next step's batch in
    # flight while this step computes (every step's copies and its loss read-back are inside the timed region).
    copy_stream = torch.cuda.Stream(device=dev)
    inflight = []

    def fetch():
        with torch.cuda.stream(copy_stream):
            lr = lr_host.to(dev, non_blocking=True)
            hr = hr_host.to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        inflight.append((lr, hr, ev))

    def e2e_step():
        if not inflight:
            fetch()
        lr, hr, ev = inflight.pop(0)
        cur = torch.cuda.current_stream()
        cur.wait_event(ev)
        lr.record_stream(cur)
        hr.record_stream(cur)
        fetch()
        losses.append(float(step(lr, hr).item()))

    e2e_step()
    ms_e2e = timed(e2e_step, args.steps)

    # ---- inference (the metric's second half): eval-mode forward of the same windows, device resident ----
    model.eval()

    def infer_step():
        with torch.no_grad():
            model(lr_dev)

    infer_step()
    ms_infer = timed(infer_step, args.steps)
    model.train()

    # ---- BASELINE.json configs[2]: x4 inference, 320x180 -> 1280x720, T=5 (clip-sharded: each rank its own clips) ----
    ms_x4, x4_batch, x4_e2e, cfg4, cfg5 = None, 16, None, None, None
    full_cfg = args.dtype == "bf16" and (args.height, args.width) == (360, 640)
    if full_cfg:
        from nerve_cl_b200.models import EnhancementConfig, EnhancementEngine
        from nerve_cl_b200.continual import EWC
        eng4 = EnhancementEngine(EnhancementConfig(frame_recovery_enabled=False, scale_factor=4, sr_num_features=args.features,
                                                   sr_num_residual_blocks=args.blocks, sr_temporal_window=2)).to(dev).eval()
        m4 = eng4.super_resolution
        m4.compute_dtype = torch.bfloat16
        lr4 = torch.rand(x4_batch, 5, 3, 180, 320, device=dev)

        def infer4_step():
            with torch.no_grad():
                m4(lr4)

        infer4_step()
        infer4_step()
        ms_x4 = timed(infer4_step, args.steps)
        del lr4
        # end to end through the public API: one 32-frame clip per step, pinned host frames -> device ->
        # EnhancementEngine.enhance_video (sliding windows batched 16 per network call) -> HR frames back in pinned host memory
        clip_frames = 32
        clip_host = torch.rand(clip_frames, 3, 180, 320).pin_memory()
        hr_host4 = [torch.empty(clip_frames, 3, 720, 1280).pin_memory() for _ in range(2)]
        d2h_stream = torch.cuda.Stream(device=dev)
        d2h_state = {"k": 0}

        def infer4_e2e():
            # the loop a user writes: clip k+1 is enhanced while clip k's HR frames travel back on a side stream
            v = clip_host.to(dev, non_blocking=True)
            hr = eng4.enhance_video(v, batch_size=x4_batch)
            done = torch.cuda.Event()
            done.record()
            with torch.cuda.stream(d2h_stream):
                d2h_stream.wait_event(done)
                hr_host4[d2h_state["k"] & 1].copy_(hr, non_blocking=True)
                hr.record_stream(d2h_stream)
            d2h_state["k"] += 1

        def run_clips():
            for _ in range(args.steps):
                infer4_e2e()
            torch.cuda.current_stream().wait_stream(d2h_stream)      # the last clip's frames are back before the clock stops

        for _ in range(max(args.warmup, 3)):       # (steady state of the caching allocator: three 354 MB HR buffers rotate)
            infer4_e2e()
        torch.cuda.synchronize()
        ms_c3 = timed(run_clips, 1)
        x4_e2e = {"metric": "sr_x4_infer_frames_per_sec", "value": clip_frames * world * args.steps / (ms_c3 / 1e3), "unit": UNIT,
                  "ms_per_clip": ms_c3 / args.steps, "frames_per_clip": clip_frames,
                  "h2d_bytes_per_step": clip_host.numel() * 4, "d2h_bytes_per_step": hr_host4[0].numel() * 4,
                  "note": "host frames -> EnhancementEngine.enhance_video (16 windows per call) -> HR frames in pinned host memory"}
        del eng4, m4

        # ---- configs[3]: continual training step = EnhancementEngine(SR-only) fwd + MSE + EWC penalty + bwd + AdamW, with
        #      8 replayed samples appended to the batch of 16 (train_continual.py's two strategies in one step) ----
        engc = EnhancementEngine(EnhancementConfig(frame_recovery_enabled=False, sr_num_features=args.features,
                                                   sr_num_residual_blocks=args.blocks)).to(dev).train()
        src = engc.super_resolution
        src.compute_dtype = torch.bfloat16
        nd.data_parallel(src)
        optc = FlatAdamW(src, lr=1e-4, weight_decay=0.0)
        ewc = EWC(src, ewc_lambda=5000.0)
        if world > 1:
            ewc.process_group = dist.group.WORLD
        ewc.register_task(0, [(lr_dev[:4], hr_dev[:4])])
        engc.train()
        lr_c = torch.cat([lr_dev, lr_dev[:8]])
        hr_c = torch.cat([hr_dev, hr_dev[:8]])
        l0 = ops.LAUNCHES[0]

        def cont_step():
            optc.zero_grad()
            loss = torch.nn.functional.mse_loss(engc(lr_c)["enhanced"], hr_c) + ewc.penalty()
            loss.backward()
            optc.step()

        cont_step()
        cont_step()
        l0 = ops.LAUNCHES[0]
        ms_c4 = timed(cont_step, args.steps)
        cfg4 = {"metric": "continual_train_frames_per_sec", "value": 24 * world * args.steps / (ms_c4 / 1e3), "unit": UNIT,
                "ms_per_step": ms_c4 / args.steps, "batch": "16 new + 8 replayed windows per GPU",
                "launches_per_step": (ops.LAUNCHES[0] - l0) / args.steps,
                "note": "EnhancementEngine(SR-only)['enhanced'] + MSE + EWC penalty (lambda 5000, online) + bwd + fused AdamW"}
        del engc, src, optc, ewc, lr_c, hr_c
        torch.cuda.empty_cache()

        # ---- configs[4]: full pipeline, FrameRecoveryNet inpainting -> x2 SR, 540p -> 1080p, host to host ----
        eng5 = EnhancementEngine(EnhancementConfig()).to(dev).eval()
        eng5.super_resolution.compute_dtype = torch.bfloat16
        eng5.frame_recovery.compute_dtype = torch.bfloat16
        n5 = 16
        clip5 = torch.rand(n5, 3, 540, 960).pin_memory()
        mask5 = torch.zeros(n5, 1, 540, 960)
        mask5[::2, :, 100:300, 200:600] = 1
        mask5 = mask5.pin_memory()
        out5 = [torch.empty(n5, 3, 1080, 1920).pin_memory() for _ in range(2)]
        d2h5 = {"k": 0}

        def pipe_step():
            # the same user loop as cfg 3: clip k + 1 is enhanced while clip k's 398 MB of HR frames travel back
            v, m = clip5.to(dev, non_blocking=True), mask5.to(dev, non_blocking=True)
            hr = eng5.enhance_video(v, m, batch_size=8)
            done = torch.cuda.Event()
            done.record()
            with torch.cuda.stream(d2h_stream):
                d2h_stream.wait_event(done)
                out5[d2h5["k"] & 1].copy_(hr, non_blocking=True)
                hr.record_stream(d2h_stream)
            d2h5["k"] += 1

        def run_pipe():
            for _ in range(max(args.steps // 2, 1)):
                pipe_step()
            torch.cuda.current_stream().wait_stream(d2h_stream)      # the last clip's frames are back before the clock stops

        for _ in range(max(args.warmup, 3)):
            pipe_step()
        torch.cuda.synchronize()
        ms_c5 = timed(run_pipe, 1)
        cfg5 = {"metric": "enhance_pipeline_frames_per_sec", "value": n5 * world * max(args.steps // 2, 1) / (ms_c5 / 1e3),
                "unit": UNIT, "ms_per_clip": ms_c5 / max(args.steps // 2, 1), "frames_per_clip": n5,
                "note": "EnhancementEngine.enhance_video: FrameRecoveryNet (4 reference frames, mask on every 2nd frame) + "
                        "SuperResolutionNet x2, 960x540 -> 1920x1080, host frames in, HR frames back in pinned host memory"}
        del eng5
        torch.cuda.empty_cache()

    # ---- the reference launcher's own configuration (train_baseline.py: 32 features / 4 dense blocks, 64x64 crops, batch
    #      16; README "~5 s/epoch estimated on CUDA" for 512 samples): launch-bound, so the step is replayed as a CUDA graph
    small = None
    if full_cfg and world == 1:
        from nerve_cl_b200.graphs import GraphedTrainStep
        torch.manual_seed(0)
        ms_small = {}
        lr_s, hr_s = synth_batch(16, 3, 64, 64, 2, 99, dev)
        for tag in ("eager", "cuda_graph"):
            m_s = SuperResolutionNet(scale_factor=2, num_features=32, num_residual_blocks=4).to(dev).train()
            m_s.compute_dtype = torch.bfloat16
            o_s = FlatAdamW(m_s, lr=1e-3, weight_decay=1e-5)
            if tag == "eager":
                def s_step():
                    o_s.zero_grad()
                    torch.nn.functional.mse_loss(m_s(lr_s), hr_s).backward()
                    o_s.step()
            else:
                g_step = GraphedTrainStep(m_s, o_s)

                def s_step():
                    g_step(lr_s, hr_s)
            for _ in range(3):
                s_step()
            ms_small[tag] = timed(s_step, 32) / 32
            del m_s, o_s
        small = {"config": "SuperResolutionNet x2, 32 feat / 4 dense blocks, 64x64 -> 128x128, T=3, B=16, bf16 (train_baseline.py defaults)",
                 "ms_per_step_eager": ms_small["eager"], "ms_per_step_cuda_graph": ms_small["cuda_graph"],
                 "epoch_512_samples_s": 32 * ms_small["cuda_graph"] / 1e3,
                 "frames_per_s_cuda_graph": 16 / (ms_small["cuda_graph"] / 1e3)}

    divergence = nd.param_divergence(model)       # max |theta_rank - theta_0| after all the steps above (must be 0)
    if rank != 0:
        return
    eager = None if args.no_gpu_eager else gpu_eager(dev, args)
    pk = peaks()
    hbm_peak = pk.get("hbm_gbs", FALLBACK_PEAKS["hbm_gbs"])
    peak = pk.get("bf16_tflops_sustained", FALLBACK_PEAKS["bf16_tflops_sustained"])

    # ---- bandwidth-bound kernels against the measured HBM copy bandwidth --------------------------------
    # (a) from the per-launch CUDA-event spans of the timed steps: algorithmic bytes per LR pixel (SURVEY.md 8d,
    #     DESIGN.md 3; e = 2 B bf16, C = 64) x pixels per launch / launch time
    e = 2 if args.dtype == "bf16" else 4
    C = args.features
    px, pxT = B * H * W, B * T * H * W
    byte_model = {                       # kernel -> (bytes per launch, launches counted per span)
        "warp_fwd": (2 * C * e + 8) * px, "warp_bwd": (3 * C * e + 16) * px, "warp_bwd_lp": (3 * C * e + 16) * px,
        "corr_fwd": (2 * C * e + 96 * e) * px, "corr_bwd": (96 * e + 4 * C * e) * px,
        "dwconv3x3_fwd": 2 * C * e * pxT, "dwconv3x3_wgrad": 2 * C * e * pxT,
        "bn_stats": C * e * pxT, "bn_relu_fwd": 2 * C * e * pxT, "bn_relu_bwd_reduce": 2 * C * e * pxT,
        "bn_relu_bwd_apply": 3 * C * e * pxT, "tfuse_fwd": ((T + 1) * C * e + 2 * T * 4) * px,
    }
    hbm_kernels = {}
    for k, nbytes in byte_model.items():
        d = ksum.get(k)
        if d and d["ms"] > 0:
            gbs = nbytes * d["launches"] / (d["ms"] / 1e3) / 1e9
            hbm_kernels[k] = {"GBs": round(gbs, 1), "frac": round(gbs / hbm_peak, 3),
                              "ms_per_launch": round(d["ms"] / d["launches"], 4)}
    # (b) the flat-buffer kernels are launch-bound at the model's 2 M parameters (24 MB: L2 resident), so their
    #     bandwidth is measured on a 2^26-element buffer as well
    n_big = 1 << 26
    th = torch.randn(n_big, device=dev)
    fi = torch.rand(n_big, device=dev)
    st = torch.randn(n_big, device=dev)
    gr = torch.randn(n_big, device=dev)
    acc = torch.zeros(1, device=dev)
    ea, eq = torch.zeros(n_big, device=dev), torch.zeros(n_big, device=dev)

    def dev_ms(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    flat = {
        "ewc_penalty_fwd": (12, lambda: ops.nv.ewc_penalty_fwd([th], fi, st, 2500.0, acc)),
        "ewc_penalty_bwd": (20, lambda: ops.nv.ewc_penalty_bwd([th], [gr], fi, st, 5000.0, None)),
        "ewc_fisher_accum": (12, lambda: ops.nv.ewc_fisher_accum(fi, [gr], [n_big], 1.0)),
        "adamw_step": (28, lambda: ops.nv.adamw_step(th, gr, ea, eq, 1e-3, 0.9, 0.999, 1e-8, 1e-5, 1, 1.0)),
    }
    for k, (bpe, fn) in flat.items():
        ms_k = dev_ms(fn)
        gbs = bpe * n_big / (ms_k / 1e3) / 1e9
        hbm_kernels[k] = {"GBs": round(gbs, 1), "frac": round(gbs / hbm_peak, 3), "ms_per_launch": round(ms_k, 4),
                          "elements": n_big}
    del th, fi, st, gr, ea, eq
    frames = B * world * args.steps
    value = frames / (ms / 1e3)
    conv_ms = sum(d["ms"] for k, d in ksum.items() if k.startswith("conv"))
    conv_flops = sum(d["flops"] for k, d in ksum.items() if k.startswith("conv"))
    conv_launches = sum(d["launches"] for k, d in ksum.items() if k.startswith("conv"))
    achieved = conv_flops / (conv_ms / 1e3) / 1e12 if conv_ms > 0 else 0.0
    peak = pk.get("bf16_tflops_sustained", FALLBACK_PEAKS["bf16_tflops_sustained"])
    # Which roofline binds each conv launch?  Per shape: t_tensor = algorithmic FLOPs / measured bf16 peak, t_hbm =
    # algorithmic bytes (input + output channels of the launch, bf16) / measured copy bandwidth; the launch cannot
    # finish before max(t_tensor, t_hbm).  Many launches of this network are 1x1 or narrow convolutions whose
    # t_hbm exceeds t_tensor, so the family's distance to the TENSOR roofline alone overstates the headroom.
    bind_ms, hbm_bound_ms, hbm_bound_launch_ms = 0.0, 0.0, 0.0
    for k, d in kdetail.items():
        if not k.startswith("conv") or d["ms"] <= 0:
            continue
        t_t = d["flops"] / (peak * 1e12) * 1e3
        t_h = d["bytes"] / (hbm_peak * 1e9) * 1e3
        bind_ms += max(t_t, t_h)
        if t_h > t_t:
            hbm_bound_ms += t_h
            hbm_bound_launch_ms += d["ms"]
    # DRAM bytes per conv-family launch from the committed ncu pass (profiles/conv_traffic.json, written by
    # scripts/summarize_profile.py from `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum` over one step)
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "conv_traffic.json")) as f:
            tj = json.load(f)
        if tj.get("workload") == workload_config(args, world).get("workload"):
            traffic, traffic_src = tj["dram_bytes_per_launch"], tj.get("source")
    except (OSError, ValueError, KeyError):
        pass
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": workload_config(args, world),
        "clocks": clocks,
        "e2e": {"value": frames / (ms_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches,
        "roofline": {
            "kernel": "dense conv family (fwd + dgrad + wgrad launches of nervecl_conv2d_*)",
            "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
            "frac": achieved / peak if peak else None, "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({pk['_source']})",
            "launches_timed": conv_launches, "avg_launch_ms": conv_ms / max(conv_launches, 1),
            "share_of_step": conv_ms / ms_spans if ms_spans else None,
            "binding": {"note": "sum over launches of max(algorithmic FLOPs / bf16 peak, algorithmic bytes / HBM peak) / measured time",
                        "frac": bind_ms / conv_ms if conv_ms else None, "lower_bound_ms_per_step": bind_ms / args.steps,
                        "measured_ms_per_step": conv_ms / args.steps,
                        "hbm_bound_share_of_conv_time": hbm_bound_launch_ms / conv_ms if conv_ms else None},
            "span_pass_ms_per_step": ms_spans / args.steps,
            "by_kind": {k: {"launches": d["launches"], "ms": round(d["ms"], 3),
                            "tflops": d["flops"] / (d["ms"] / 1e3) / 1e12 if d["ms"] > 0 else 0.0}
                        for k, d in ksum.items() if k.startswith("conv")},
        },
        "infer": {"metric": "sr_x2_infer_frames_per_sec", "value": frames / (ms_infer / 1e3), "unit": UNIT,
                  "ms_per_batch": ms_infer / args.steps, "note": "eval-mode forward only, same windows, inputs in HBM"},
        "infer_x4": ({"metric": "sr_x4_infer_frames_per_sec", "value": x4_batch * world * args.steps / (ms_x4 / 1e3),
                      "unit": UNIT, "ms_per_batch": ms_x4 / args.steps,
                      "note": "BASELINE configs[2]: scale 4, T=5, 320x180 -> 1280x720, B=16/GPU, eval forward, inputs in HBM"}
                     if ms_x4 else None),
        "infer_x4_e2e": x4_e2e,
        "continual_train": cfg4,
        "enhance_pipeline": cfg5,
        "launcher_config": small,
        "hbm_kernels": {"peak_GBs": hbm_peak, "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({pk['_source']})",
                        "kernels": hbm_kernels},
        "loss_last": losses[-1] if losses else None,
        "dp_param_divergence": divergence,
        "gpu_eager": eager,
        "timing_retries": timing_notes,
        "value_passes_ms_per_step": [v / args.steps for v in value_passes],
        "env_overrides": sorted(k for k in os.environ if k.startswith("NERVECL_")),
        "other_kernels_ms_per_step": {k: round(d["ms"] / args.steps, 3) for k, d in
                                      sorted(ksum.items(), key=lambda kv: -kv[1]["ms"]) if not k.startswith("conv")},
    }
    if not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        arm = CpuArm(args.features, args.blocks)
        rows = pick_cpu_sample(arm, args, args.cpu_budget)
        t = arm.step_seconds(1, rows, W)
        what = "the reference's own SuperResolutionNet (oracle/_ref)" if arm.kind == "reference" else "oracle port"
        line["cpu_baseline"] = {
            "value": (rows / H) / t, "unit": UNIT, "cores": cores, "kind": arm.kind, "cpu": lscpu_model(),
            "sample": (f"{what} (fp32 ATen-CPU, {torch.get_num_threads()} threads): 1 clip x {rows}/{H} rows of "
                       f"the {W}x{H} window, one fwd+MSE+bwd step = {t:.1f} s, scaled by pixel count"),
            "cfg1": cpu_cfg1(arm),
        }
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)
    try:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


if __name__ == "__main__":
    main()
