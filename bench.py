#!/usr/bin/env python
"""bench.py -- ResNet-50 4-bit GPFQ (bs=256) on N B200s: wall time and weights*samples/s.

    python bench.py --gpus 1 --steps 3 --warmup 3                      # this repo's CUDA path
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W      # neuron-sharded over N GPUs
    python bench.py --impl reference --steps K --warmup W               # CPU arm (oracle port of the reference)

One "step" = one full ``QuantizeNeuralNet(...).quantize_network()`` over the 54 layers of a
random-init torchvision ResNet-50 with synthetic Gaussian images (a fresh batch per layer, as
the reference draws them), i.e. the region the reference times at src/main.py:120-122.
Units: one (weight, calibration-row) pair, sum_layers N*d*m.

Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes
import gc
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

# dram__bytes_read.sum + dram__bytes_write.sum of one sweep_kernel<4> launch, from the committed ncu --set full capture
NCU_SWEEP_TRAFFIC = {"dram_bytes_per_launch": 185.2e6, "algorithmic_bytes_same_launch": 122.0e6,
                     "launch": "one 32-feature block of a (N=512, d=256, m=50432) layer, 214 us",
                     "source": "profiles/r01_sweep_kernel_ncu_full_summary.txt"}

# the same for one bn_act_kernel launch (filled from profiles/r01_bn_act_kernel_ncu_full_summary.txt)
NCU_BN_ACT_TRAFFIC = {"dram_bytes_per_launch": 2440.1e6, "algorithmic_bytes_same_launch": 2466.3e6,
                      "launch": "BatchNorm + residual + ReLU of a (256, 256, 56, 56) activation",
                      "source": "profiles/r01_bn_act_kernel_ncu_full_summary.txt"}

# dram__bytes_read.sum + dram__bytes_write.sum of one conv1x1_tc_kernel launch, from the committed ncu --set full capture
NCU_CONV_TRAFFIC = {"dram_bytes_per_launch": 2104.2e6, "algorithmic_bytes_same_launch": 1849.7e6,
                    "launch": "1x1 convolution 64 -> 256 + BatchNorm + residual + ReLU of a (256, 64, 56, 56) activation, 443 us",
                    "source": "profiles/r02b_conv_64_256_res_ncu_full_summary.txt"}

L2_PEAK_GBS = 8300.0     # L2-resident read bandwidth measured by tools/microbench.cu on this pool's B200 (r01)

METRIC = "resnet50_4bit_gpfq_weights_samples_per_s"   # the metric name follows --model/--bits when they differ
UNIT = "weights*samples/s"


# ----------------------------------------------------------------------------- workload
def build_model(name):
    import torchvision
    torch.manual_seed(0)
    return getattr(torchvision.models, name)(weights=None).eval()


def layer_shapes(model, batch, retain, image=224):
    """(N, d, m, groups) of every quantizable layer, walked in the reference's layer order (CUDA arm only: it
    imports the product package)."""
    from quantized_neural_nets_b200.utils import extract_layers
    layers = []
    extract_layers(model, layers)
    spatial = {}

    def hook(mod, args, out):
        spatial[mod] = tuple(args[0].shape)

    handles = [l.register_forward_hook(hook) for l in layers]
    with torch.no_grad():
        model(torch.zeros(1, 3, image, image, device=next(model.parameters()).device))
    for h in handles:
        h.remove()
    shapes = []
    for l in layers:
        if isinstance(l, torch.nn.Linear):
            shapes.append((l.out_features, l.in_features, batch, 1))
        else:
            _, C, H, W = spatial[l]
            kh, kw = l.kernel_size
            Lh = (H + 2 * l.padding[0] - l.dilation[0] * (kh - 1) - 1) // kh + 1
            Lw = (W + 2 * l.padding[1] - l.dilation[1] * (kw - 1) - 1) // kw + 1
            L = Lh * Lw
            keep = int(retain * L + 1 if retain != 1 else retain * L)
            shapes.append((l.out_channels, C // l.groups * kh * kw, batch * keep, l.groups))
    return shapes


class BatchPool:
    """Synthetic ImageNet-shaped loader: ``pool`` distinct Gaussian batches cycled, so consecutive
    layers always see different batches (the reference draws a fresh batch per layer)."""

    def __init__(self, batches):
        self.batches = batches

    def __iter__(self):
        i = 0
        while True:
            yield self.batches[i % len(self.batches)], None
            i += 1


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index, interval_ms=500):
        self.index = index
        self.interval_ms = interval_ms
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", str(self.interval_ms), "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "sm_min_mhz": min(sm) if sm else None, "sm_mhz_samples": sm[:64],
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU arm
def run_reference_arm(args):
    """``--impl reference``: the reference's own CPU implementation of the path on the box's host cores (the
    UNMODIFIED reference from oracle/_ref when oracle/make_ref.py has vendored it, else the oracle port), every step
    a bounded sample of the workload (oracle/cpu_baseline.py).  Rank 0 alone works; nothing of the product is
    imported."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    os.environ.setdefault("TQDM_DISABLE", "1")
    from oracle import cpu_baseline as cb
    cores = cb.host_threads()
    _, _, extract_layers, kind = cb.reference_modules()
    shapes = cb.layer_shapes(cb.build_model(args.model), args.batch, args.retain, extract_layers)
    units = float(sum(N * d * m for (N, d, m, g) in shapes))
    for _ in range(args.warmup):       # cheap warm-up samples: short classes, an 8-image forward scaled to the batch
        cb.sample_step(args.model, args.batch, args.retain, args.bits, shapes, seconds_per_class=0.05, forward_batch=8)
    steps = [cb.sample_step(args.model, args.batch, args.retain, args.bits, shapes, seconds_per_class=args.cpu_class_seconds)
             for _ in range(args.steps)]
    seconds = sum(s["seconds"] for s in steps) / len(steps)
    value = units / seconds
    validation = cb.validate(batch=args.batch, retain=args.retain) if args.validate_cpu else None
    base = {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": cb.describe(steps[-1], cores, len(shapes), args.batch),
            "sampling_ms_per_step": 1e3 * sum(s["spent"] for s in steps) / len(steps)}
    if validation is not None:
        base["validation"] = validation
    print(json.dumps({
        "impl": "reference", "metric": metric_name(args), "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * seconds, "ms_per_step_note": "implied by value (extrapolated from the bounded sample), "
        "not the sampling time", "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args), "cpu_baseline": base,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(args):
    return {"workload": f"{args.model} {args.bits}-bit GPFQ, bs={args.batch}, retain_rate={args.retain}, scalar=1.16"
                        + (f", reg={args.reg} lamb={args.lamb}" if args.reg else "")
                        + (" (BASELINE.json configs[3])" if args.model == "resnet50" else "")
                        + "; one step = quantize_network() over all layers",
            "global_batch": args.batch, "image": "3x224x224 Gaussian", "weights": "random init (torch.manual_seed(0))",
            "l2": "per-step inputs (8.3 GB of images, up to 400 MB of layer inputs) exceed the 126 MB L2",
            "parallelism": "single GPU" if args.gpus == 1 else
                           f"neuron-sharded x{args.gpus}, {args.forward} calibration forward, "
                           + ("1 all-gather (Q) per layer" if args.forward == "replicated"
                              else "2 all-gathers (layer inputs, Q) per layer")}


# ----------------------------------------------------------------------------- CUDA arm
def run_cuda_arm(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import quantized_neural_nets_b200 as qb
    from quantized_neural_nets_b200 import _lib

    torch.backends.cudnn.allow_tf32 = False           # fp32 calibration forward
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = True

    model = build_model(args.model).to(dev)
    shapes = layer_shapes(model, args.batch, args.retain)
    units = float(sum(N * d * m for (N, d, m, g) in shapes))
    gen = torch.Generator(device=dev).manual_seed(1)
    dev_pool = [torch.randn(args.batch, 3, 224, 224, device=dev, generator=gen) for _ in range(args.pool)]
    host_pool = [b.cpu().pin_memory() for b in dev_pool]
    img_bytes = dev_pool[0].numel() * 4

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step(pool, read_back, profile=False, forward=None, calibration=None, fuse=None):
        forward = forward or args.forward
        np.random.seed(0)
        qnn = qb.QuantizeNeuralNet(model, args.model, args.batch, BatchPool(pool), args.bits, args.bits, [],
                                   1.16, 1.16, 1, 1, args.reg, args.lamb, args.retain, False, dev, profile=profile,
                                   shard_forward=(forward == "sharded" and world > 1),
                                   solver=None if args.solver == "direct" else args.solver,
                                   calibration=calibration or args.calibration,
                                   fuse_forward=args.fuse_forward if fuse is None else fuse,
                                   pointwise_gemm=args.pointwise_gemm if fuse is None else fuse)
        # every step builds two fresh network copies and two torch.fx graphs; collect that garbage HERE, outside the timed
        # region, so that a full (generation-2) collection of Python's cyclic GC does not land inside it (seen as one step
        # in three taking +240 ms with the GPU idle)
        gc.collect()
        barrier()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        qnn.quantize_network()
        d2h = 0
        if read_back:
            errs = torch.stack([torch.stack((e, r)) for (_, e, r) in qnn.layer_log]).cpu()
            d2h = errs.numel() * 4
            # the reference's relative error of a grouped layer is 0/0 = nan when a group's inputs are all zero
            # (dead channels of a random-init depthwise network); anything else must be finite
            assert torch.isfinite(errs[:, 0]).all() and (args.model != "resnet50" or torch.isfinite(errs).all())
        end.record()
        barrier()
        return start.elapsed_time(end), qnn, d2h

    STEP_LOG = []      # every timed step of this rank, in order (diagnostics: which loop a slow step belongs to)
    ALLOC_LOG = []     # cumulative (cudaMalloc calls, cudaFree calls, allocation retries) of the caching allocator after each

    def timed(pool, read_back, sampler=None, forward=None, warmup=None, calibration=None, fuse=None):
        for _ in range(args.warmup if warmup is None else warmup):
            one_step(pool, read_back, forward=forward, calibration=calibration, fuse=fuse)
        if sampler:
            sampler.start()
        before = _lib.launch_count()
        total = 0.0
        d2h = 0
        for _ in range(args.steps):
            ms, qnn, d2h = one_step(pool, read_back, forward=forward, calibration=calibration, fuse=fuse)
            total += ms
            st = torch.cuda.memory_stats(dev)
            STEP_LOG.append(round(ms, 1))
            ALLOC_LOG.append((st.get("num_device_alloc", 0), st.get("num_device_free", 0), st.get("num_alloc_retries", 0)))
        launches = _lib.launch_count() - before
        clocks = sampler.stop() if sampler else None
        t = torch.tensor([total], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), launches, clocks, qnn, d2h

    if args.profile_one_step:
        # for `ncu --profile-from-start off`: warm up (autotuning included), then expose exactly ONE step to the
        # profiler.  Nothing measured under a profiler is a bench value, so nothing is printed.
        for _ in range(args.warmup):
            one_step(dev_pool, False)
        torch.cuda.profiler.start()
        one_step(dev_pool, False)
        torch.cuda.profiler.stop()
        if world > 1:
            dist.destroy_process_group()
        return None
    total_ms, launches, clocks, qnn, _ = timed(dev_pool, False, ClockSampler(local, 500 if world < 4 else 200) if rank == 0 else None)
    value_steps = list(STEP_LOG)
    e2e_ms, _, _, qnn, d2h = timed(host_pool, True)
    e2e_steps = STEP_LOG[len(value_steps):]
    # the box's pinned host -> device rate (one batch): the e2e step copies `layers` batches, one per layer, prefetched one
    # layer ahead; the first layers' prefix passes are shorter than a copy, so e2e - value grows as this rate drops
    a_ev, b_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a_ev.record()
    for hb in host_pool[:4]:
        hb.to(dev, non_blocking=True)
    b_ev.record()
    torch.cuda.synchronize()
    h2d_gbs = 4 * img_bytes / (a_ev.elapsed_time(b_ev) * 1e-3) / 1e9
    other_ms = None
    if world > 1:   # the other multi-GPU forward mode, for the record (same K, one warm-up)
        other = "replicated" if args.forward == "sharded" else "sharded"
        other_ms, _, _, _, _ = timed(dev_pool, False, forward=other, warmup=1)
    n_layers = len(qnn.layer_log)
    rel = [float(r) for (_, _, r) in qnn.layer_log]
    unfused_ms = None
    if args.fuse_forward and world == 1:   # the same step through PyTorch's own BatchNorm / add / ReLU kernels, for the record
        unfused_ms, _, _, _, _ = timed(dev_pool, False, fuse=False, warmup=1)
    reuse_ms = reuse_e2e_ms = None
    if args.calibration == "fresh":   # the O(L) single-batch schedule, for the record
        reuse_ms, _, _, _, _ = timed(dev_pool, False, calibration="reuse")
        reuse_e2e_ms, _, _, _, _ = timed(host_pool, True, calibration="reuse", warmup=1)

    # one extra, untimed step with per-launch CUDA events around the dominant kernel (the sweep)
    _lib.profile_begin()
    step_ms, _, _ = one_step(dev_pool, False)
    prof = _lib.profile_end()
    phases = per_layer = None
    for _ in range(2):      # phase split: the steadier of two instrumented steps (a one-off stall would skew a single one)
        _, qprof, _ = one_step(dev_pool, False, profile=True)
        ph, pl = qprof.phase_times_ms()
        if phases is None or sum(ph.values()) < sum(phases.values()):
            phases, per_layer = ph, pl

    out = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        fp32_peak = 148 * 128 * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6   # fp32 instr/s at max clock
        sweep_s = prof["sweep_ms"] * 1e-3
        achieved = prof["sweep_bytes"] / sweep_s / 1e9 if sweep_s > 0 else 0.0
        ms_per_step = total_ms / args.steps
        out = {
            "metric": metric_name(args), "value": units / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "wall_time_s": ms_per_step * 1e-3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": f"synthetic ({args.pool} Gaussian image batches cycled, random-init weights)",
            "config": workload_config(args),
            "units_per_step": units, "layers": n_layers,
            "e2e": {"value": units / (e2e_ms / args.steps * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": n_layers * img_bytes, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / args.steps, "steps_ms": e2e_steps, "pinned_h2d_gbs": round(h2d_gbs, 1)},
            "steps_ms": value_steps,
            "allocator_after_each_step": ALLOC_LOG[:len(value_steps) + len(e2e_steps)],
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": None,      # filled below: the kernel of this library with the largest share of the step
            "direct_kernel_roofline": {
                "kernel": "gpfq::sweep_kernel", "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak,
                "traffic": NCU_SWEEP_TRAFFIC["dram_bytes_per_launch"], "traffic_detail": NCU_SWEEP_TRAFFIC,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                "launches_per_step": prof["sweep_launches"],
                "avg_launch_us": 1e3 * prof["sweep_ms"] / max(1, prof["sweep_launches"]),
                "kernel_ms_per_step": prof["sweep_ms"], "share_of_step": prof["sweep_ms"] / step_ms,
                "l2": {"note": "north_star asks for the direct kernel against the L2 bandwidth roofline: algorithmic L2 -> SM "
                               "bytes (SURVEY.md 8d: the HBM bytes plus the X / Xq block tiles once per further neuron "
                               "tile of the grid) over the same CUDA-event time; peak = L2-resident read bandwidth "
                               "measured on this pool's B200 by tools/microbench.cu (r01)",
                       "achieved": prof["sweep"]["aux"] / sweep_s / 1e9 if sweep_s > 0 else 0.0, "peak": L2_PEAK_GBS,
                       "unit": "GB/s", "frac": (prof["sweep"]["aux"] / sweep_s / 1e9) / L2_PEAK_GBS if sweep_s > 0 else 0.0},
                "fp32": {"note": "the sweep is fp32-issue bound by design: 5 separately rounded fp32 instructions per "
                                 "(neuron, sample, feature); at the fp32 issue peak it would move 8 bytes of U per 160 "
                                 "instructions = 1.9 TB/s, so neither HBM nor L2 can be its bound",
                         "achieved_ginstr_s": prof["sweep_fp32_instr"] / sweep_s / 1e9 if sweep_s > 0 else 0.0,
                         "peak_ginstr_s": fp32_peak / 1e9,
                         "frac": (prof["sweep_fp32_instr"] / sweep_s) / fp32_peak if sweep_s > 0 else 0.0},
            },
            "resident_kernel": {
                "note": "single-launch structure of the direct solver for launch-bound layers (U in shared memory, "
                        "cluster-split columns); latency-chain bound, so it is reported by time, not against HBM",
                "launches_per_step": prof["resident_launches"], "kernel_ms_per_step": prof["resident_ms"],
                "share_of_step": prof["resident_ms"] / step_ms,
                "fp32_frac": (prof["resident_fp32_instr"] / (prof["resident_ms"] * 1e-3)) / fp32_peak
                if prof["resident_ms"] > 0 else 0.0},
            "gram_kernel_roofline": gram_roofline(prof, peaks, step_ms),
            "conv_kernel": conv_summary(prof, hbm_peak, peaks, step_ms),
            "parity": {"min_level_agreement_of_gated_layers": sa_min_agreement(),
                       "note": "per LAYER (not per shape): a Gram variant -- and, sharded, the Gram-reduce mode -- is used "
                               "for a layer only if it reproduced >= 99.9 % of the direct solver's levels on that layer's own "
                               "data during warm-up; the direct solver itself is pinned against the reference by the "
                               "bit-exact golden tests and the teacher-forced network tests (tests/)"},
            "rel_err_mean": sum(r for r in rel if r == r) / max(1, sum(1 for r in rel if r == r)),
            "rel_err_nan_layers": sum(1 for r in rel if r != r),
            "forward_mode": args.forward if world > 1 else "single GPU",
            "calibration": args.calibration,
            "fuse_forward": args.fuse_forward,
            "pointwise_gemm": args.pointwise_gemm,
            "solver": args.solver,
            "solver_choices": solver_choices(),
            "phase_ms_per_step": {k: round(v, 2) for k, v in phases.items()},
            "solve_ms_per_layer": [round(per_layer[i].get("solve", 0.0), 3) for i in sorted(per_layer)],
        }
        if prof["conv"]["ms"] > max(prof["sweep_ms"], prof["resident_ms"], prof["bn_act_ms"]):
            c = prof["conv"]
            gbs = c["bytes"] / (c["ms"] * 1e-3) / 1e9
            out["roofline"] = {
                "kernel": "gpfq::conv1x1_tc_kernel", "bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                "frac": gbs / hbm_peak, "traffic": NCU_CONV_TRAFFIC["dram_bytes_per_launch"], "traffic_detail": NCU_CONV_TRAFFIC,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                "launches_per_step": c["launches"], "avg_launch_us": 1e3 * c["ms"] / max(1, c["launches"]),
                "kernel_ms_per_step": c["ms"], "share_of_step": c["ms"] / step_ms,
                "tensor": out["conv_kernel"]["tensor"],
                "note": "the kernel of this library with the largest share of the step: the calibration forward's "
                        "convolutions (tcgen05 split-TF32 GEMM + fused BatchNorm / residual / ReLU epilogue); algorithmic "
                        "bytes = read the activation (or patch matrix) once (+ residual), write the output once, summed "
                        "over all its launches of one step.  The GPFQ kernels are in direct_kernel_roofline / "
                        "resident_kernel / gram_kernel_roofline."}
        elif prof["bn_act_ms"] > max(prof["sweep_ms"], prof["resident_ms"]):
            bn_gbs = prof["bn_act_bytes"] / (prof["bn_act_ms"] * 1e-3) / 1e9
            out["roofline"] = {
                "kernel": "gpfq::bn_act_kernel", "bound": "hbm", "achieved": bn_gbs, "peak": hbm_peak, "unit": "GB/s",
                "frac": bn_gbs / hbm_peak,
                "traffic": NCU_BN_ACT_TRAFFIC["dram_bytes_per_launch"], "traffic_detail": NCU_BN_ACT_TRAFFIC,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                "launches_per_step": prof["bn_act_launches"],
                "avg_launch_us": 1e3 * prof["bn_act_ms"] / max(1, prof["bn_act_launches"]),
                "kernel_ms_per_step": prof["bn_act_ms"], "share_of_step": prof["bn_act_ms"] / step_ms,
                "note": "fused inference BatchNorm (+ residual add) (+ ReLU) of the calibration forward: the kernel of "
                        "this library with the largest share of the step; algorithmic bytes = read x (+ residual) + "
                        "write out.  The GPFQ direct kernel is in direct_kernel_roofline / resident_kernel."}
        else:
            out["roofline"] = out["direct_kernel_roofline"]
        if unfused_ms is not None:
            out["unfused_forward"] = {"note": "fuse_forward=False, pointwise_gemm=False: cuDNN inference batch norm, "
                                              "separate add / ReLU kernels and cuDNN 1x1 convolutions in the "
                                              "calibration forward",
                                      "ms_per_step": unfused_ms / args.steps,
                                      "value": units / (unfused_ms / args.steps * 1e-3)}
        if reuse_ms is not None:
            out["reuse_calibration"] = {
                "note": "calibration='reuse': ONE batch calibrates all layers (one analog pass + one quantizing pass "
                        "instead of two prefix passes per layer); NOT the reference's fresh-batch-per-layer schedule, "
                        "so it is reported beside the headline, never as it",
                "ms_per_step": reuse_ms / args.steps, "value": units / (reuse_ms / args.steps * 1e-3),
                "e2e_ms_per_step": reuse_e2e_ms / args.steps, "e2e_value": units / (reuse_e2e_ms / args.steps * 1e-3),
                "h2d_bytes_per_step": img_bytes // (world if args.forward == "sharded" else 1)}
        if other_ms is not None:
            out["other_forward_mode"] = {"mode": other, "ms_per_step": other_ms / args.steps,
                                         "value": units / (other_ms / args.steps * 1e-3)}
        if world == 1 and args.all_configs and args.model == "resnet50" and args.bits == 4:
            out["all_configs"] = measure_other_configs(args, dev)
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline_leg(args, shapes, units)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return out


def sa_min_agreement():
    from quantized_neural_nets_b200 import step_algorithm as sa
    return round(sa.min_gated_agreement(), 6)


def gram_roofline(prof, peaks, step_ms):
    """tensor-pipe view of the Gram solver's GEMM kernel (north_star: 'tensor-pipe utilisation for the Gram GEMMs')."""
    g, gp = prof["gram_tc"], prof["gram_path"]
    tf32_peak = 0.5 * float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1400.0)))
    s = g["ms"] * 1e-3
    alg = g["flops"] / s / 1e12 if s > 0 else 0.0
    return {"kernel": "gpfq::gram_tc_kernel", "bound": "tensor", "unit": "TFLOP/s",
            "achieved_algorithmic": alg, "achieved_issued": 3.0 * alg, "peak": tf32_peak,
            "peak_source": "half of MEASURED_PEAKS.json bf16_tflops_sustained (dense TF32 = bf16 / 2)" if peaks else "fallback",
            "frac": 3.0 * alg / tf32_peak if tf32_peak > 0 else 0.0,
            "note": "split-TF32: three MMAs per product, so issued = 3 x algorithmic; flops = the 128 x 128 tile products formed",
            "launches_per_step": g["launches"], "kernel_ms_per_step": g["ms"], "share_of_step": g["ms"] / step_ms,
            "gram_path_kernel": {"launches_per_step": gp["launches"], "kernel_ms_per_step": gp["ms"],
                                 "fp64_tflops": 2.0 * gp["flops"] / (gp["ms"] * 1e-3) / 1e12 if gp["ms"] > 0 else 0.0},
            "recur_kernel": {"launches_per_step": prof["recur"]["launches"], "kernel_ms_per_step": prof["recur"]["ms"]}}


def conv_summary(prof, hbm_peak, peaks, step_ms):
    c = prof["conv"]
    s = c["ms"] * 1e-3
    tf32_peak = 0.5 * float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1400.0)))
    alg = c["flops"] / s / 1e12 if s > 0 else 0.0
    return {"kernel": "gpfq::conv1x1_tc_kernel", "launches_per_step": c["launches"], "kernel_ms_per_step": c["ms"],
            "share_of_step": c["ms"] / step_ms, "hbm_gbs": c["bytes"] / s / 1e9 if s > 0 else 0.0,
            "hbm_frac": (c["bytes"] / s / 1e9) / hbm_peak if s > 0 else 0.0,
            "tensor": {"achieved_algorithmic_tflops": alg, "achieved_issued_tflops": 3.0 * alg, "peak_tflops": tf32_peak,
                       "frac": 3.0 * alg / tf32_peak if tf32_peak > 0 else 0.0}}


def measure_other_configs(args, dev):
    """BASELINE.json's other configurations on one GPU, driver-observed: {name: summary}.  Each: 2 warm-up steps (per-layer
    solver choice behind the parity gate) + 2 timed steps of quantize_network() with the bench's settings, and one
    instrumented step for the forward / solve split."""
    import quantized_neural_nets_b200 as qb
    from quantized_neural_nets_b200 import _lib, step_algorithm as sa
    out = {}
    for tag, name, bits, reg, lamb in (("alexnet_4bit", "alexnet", 4, None, 0.1), ("resnet18_4bit", "resnet18", 4, None, 0.1),
                                       ("vgg16_4bit_L1", "vgg16", 4, "L1", 0.1), ("resnet50_3bit", "resnet50", 3, None, 0.1)):
        model = build_model(name).to(dev)
        shapes = layer_shapes(model, args.batch, args.retain)
        units = float(sum(N * d * m for (N, d, m, g) in shapes))
        gen = torch.Generator(device=dev).manual_seed(1)
        pool = [torch.randn(args.batch, 3, 224, 224, device=dev, generator=gen) for _ in range(2)]
        log0 = len(sa.AUTO_LOG)

        def step(profile=False):
            np.random.seed(0)
            qnn = qb.QuantizeNeuralNet(model, name, args.batch, BatchPool(pool), bits, bits, [], 1.16, 1.16, 1, 1, reg, lamb,
                                       args.retain, False, dev, profile=profile,
                                       solver=None if args.solver == "direct" else args.solver,
                                       fuse_forward=args.fuse_forward, pointwise_gemm=args.pointwise_gemm)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            qnn.quantize_network()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b), qnn

        for _ in range(2):
            step()
        ms = sum(step()[0] for _ in range(2)) / 2
        _, qprof = step(profile=True)
        phases, per_layer = qprof.phase_times_ms()
        rel = [float(r) for (_, _, r) in qprof.layer_log]
        picked = [ag for (k, tm, ag, ch) in sa.AUTO_LOG[log0:] if ch not in (_lib.SOLVER_DIRECT, "groups_loop", "gather_inputs")]
        out[tag] = {"workload": f"{name} {bits}-bit bs={args.batch} retain={args.retain}" + (f" reg={reg} lamb={lamb}" if reg else ""),
                    "units_per_step": units, "layers": len(shapes), "ms_per_step": ms, "value": units / (ms * 1e-3),
                    "phase_ms": {k: round(v, 2) for k, v in phases.items()},
                    "rel_err_mean": sum(r for r in rel if r == r) / max(1, sum(1 for r in rel if r == r)),
                    "gram_layers": len(picked), "min_level_agreement_of_gated_layers": round(min(picked), 6) if picked else 1.0}
        if name == "vgg16":
            out[tag]["fc6_variants"] = fc6_variants(dev, per_layer[13].get("solve", 0.0))
        del model, pool, qprof
        torch.cuda.empty_cache()
    return out


def fc6_variants(dev, direct_ms):
    """BASELINE.json configs[4] names VGG-16 fc6 (4096 x 25088, m = 256) as the Gram / tcgen05 showcase, SURVEY.md 8a
    asks that both variants be timed on it and the choice reported honestly.  m << d is the opposite of the Gram regime:
    the three 25088 x 25088 fp64 Gram matrices are 15.1 GB and the recurrence needs about 2.5 N d^2 = 6.4e12 fp64 FMAs.
    Timed here: the direct solver on the layer (from the instrumented step) and the Gram FORMATION alone (tcgen05,
    gpfq_gram_f32); the recurrence kernel keeps a neuron's (w, q) rows in shared memory and supports d <= 3104, so its
    cost is given as a lower bound from the fp64 rate it reaches on the layers it does run."""
    from quantized_neural_nets_b200 import _lib
    lib, launch = _lib.lib, _lib.launch
    d, m, N = 25088, 256, 4096
    out = {"direct_ms": round(direct_ms, 3), "chosen": "direct"}
    try:
        g = torch.Generator(device=dev).manual_seed(5)
        X = torch.relu(torch.randn(d, m, device=dev, generator=g))
        ldg = (d + 63) // 64 * 64
        grams = torch.empty((3, ldg, ldg), dtype=torch.float64, device=dev)
        nbytes = lib.gpfq_gram_workspace_bytes(_lib.SOLVER_GRAM, d, m)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        for _ in range(2):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            launch(lib.gpfq_gram_f32, _lib.SOLVER_GRAM, X, X, m, d, m, grams[0], grams[1], grams[2], ws, nbytes)
            b.record()
            torch.cuda.synchronize()
        out["gram_formation_ms"] = round(a.elapsed_time(b), 3)
        out["gram_matrices_gb"] = round(3 * ldg * ldg * 8 / 1e9, 1)
        out["gram_recurrence_lower_bound_ms"] = round(2.5 * N * float(d) * d / 9.0e12 * 1e3, 1)
        out["note"] = ("Gram recurrence bound = 2.5 N d^2 fp64 FMAs at 9 TFMA/s (the fp64 FMA rate of B200 measured by "
                       "tools/microbench.cu is 18 TFLOP/s); the direct solver wins by more than an order of magnitude, as "
                       "m << d predicts")
        del grams, ws, X
        torch.cuda.empty_cache()
    except Exception as exc:          # out of memory on a smaller part: the comparison is informative, not required
        out["gram_formation_error"] = str(exc)[:200]
    return out


def full_size_spot_check(kept):
    """Parity at the bench's FULL layer sizes (bs = 256: m up to 200,960): the (W, X) sub-problems the CPU arm has just
    run through the reference -- the first k features of one layer shape per class -- are handed to the CUDA solvers and
    the levels compared.  GPFQ decides feature t from features <= t only, so the first k columns are a complete problem."""
    from quantized_neural_nets_b200 import _lib, step_algorithm as sa
    dev = torch.device("cuda", torch.cuda.current_device())
    rows, worst = [], 1.0
    for c in kept:
        N, d, m = c["shape"]
        W, X, K = c["W"].to(dev), c["X"].to(dev), c["K"]
        delta = c["delta"].to(dev)
        lv_ref = torch.round(c["Q"].double() / float(c["delta"])).to(torch.int32)
        rec = {"shape": f"{N}x{d}x{m}", "features": c["k"]}
        for name, solver in (("direct", _lib.SOLVER_DIRECT), ("gram_tcgen05", _lib.SOLVER_GRAM)):
            if solver == _lib.SOLVER_GRAM and not sa.gram_eligible(N, c["k"], m):
                continue
            Q, _, _ = sa.quantize_layer_impl(W, X, X, m, 1.16 / K, K, 1, None, 0.1, 1, False, dev, solver=solver,
                                             return_partials=True, delta=delta)
            lv = torch.round(Q.double().cpu() / float(c["delta"])).to(torch.int32)
            rec[name] = round(float((lv == lv_ref).float().mean()), 6)
            worst = min(worst, rec[name])
        rows.append(rec)
    return {"what": "level agreement of the CUDA solvers with the reference's own decisions on the first k features of "
                    "full-size (bs = 256) layer shapes, same (W, X)", "min": worst, "per_shape": rows}


def cpu_baseline_leg(args, shapes, units):
    """``cpu_baseline`` of the CUDA arm's line (rank 0, N = 1): one bounded sample of the reference on the host cores
    (oracle/cpu_baseline.py) and, unless --no-validate-cpu, the measured-vs-extrapolated check against one complete
    real reference run (AlexNet)."""
    from oracle import cpu_baseline as cb
    cores = cb.host_threads()
    kept = []
    step = cb.sample_step(args.model, args.batch, args.retain, args.bits, shapes, seconds_per_class=args.cpu_class_seconds,
                          keep=kept)
    out = {"value": units / step["seconds"], "unit": UNIT, "cores": cores, "kind": step["kind"],
           "solver_only_value": units / step["solver_s"], "sample": cb.describe(step, cores, len(shapes), args.batch)}
    out["full_size_spot_check"] = full_size_spot_check(kept)
    if args.validate_cpu:
        out["validation"] = cb.validate(batch=args.batch, retain=args.retain)
    return out


def solver_choices():
    """{"N x d x m": solver name} for the shapes the autotuner timed (rank 0's slices)."""
    from quantized_neural_nets_b200 import step_algorithm as sa
    names = {0: "direct", 1: "gram_tcgen05", 2: "gram_f64"}

    def label(k):     # (rows, d, m, mode) for ungrouped layers, ("<groups>g", N, d_group, m) for grouped ones
        return "x".join(str(v) for v in (k if isinstance(k[0], str) else k[:3]))

    return {label(k): {"chosen": names.get(ch, ch), "agree": round(ag, 6),
                       "ms": {names.get(s, s): round(t, 3) for s, t in tm.items()}}
            for (k, tm, ag, ch) in sa.AUTO_LOG}


def metric_name(args):
    return f"{args.model}_{args.bits}bit_gpfq_weights_samples_per_s"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--model", default="resnet50")
    ap.add_argument("--bits", type=int, default=4)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--retain", type=float, default=0.25)
    ap.add_argument("--pool", type=int, default=8, help="distinct synthetic image batches cycled by the loader")
    ap.add_argument("--forward", default="sharded", choices=["replicated", "sharded"],
                    help="multi-GPU only: split each calibration batch over the ranks and all-gather the layer inputs "
                         "(default; removes the replicated-forward Amdahl term), or replicate the calibration forward "
                         "on every rank (BASELINE.json's sketch; Q bit-identical to the single-GPU run)")
    ap.add_argument("--reg", default=None, choices=[None, "L0", "L1"], help="sparse-mode regulariser (main.py -reg)")
    ap.add_argument("--lamb", type=float, default=0.1, help="regularisation strength (main.py -l)")
    ap.add_argument("--solver", default="auto", choices=["auto", "direct"],
                    help="auto: per layer, direct vs Gram (tcgen05 / fp64) picked from measured time behind a 99.9 %% "
                         "level-agreement gate during warm-up; direct: blocked direct solver everywhere")
    ap.add_argument("--calibration", default="fresh", choices=["fresh", "reuse"],
                    help="fresh: a new batch and two prefix forward passes per layer (the reference's schedule, the "
                         "headline); reuse: one batch for all layers, two network passes in total")
    ap.add_argument("--no-fuse-forward", dest="fuse_forward", action="store_false",
                    help="run the calibration forward through PyTorch's own BatchNorm / add / ReLU kernels instead of the "
                         "fused elementwise kernel (forward_fusion.py)")
    ap.add_argument("--no-pointwise-gemm", dest="pointwise_gemm", action="store_false",
                    help="keep cuDNN for the stride-1 1x1 convolutions of the calibration forward instead of one "
                         "strided-batched cuBLAS SGEMM each (gpfq_conv1x1_f32)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-all-configs", dest="all_configs", action="store_false",
                    help="N = 1 only: skip the summary block of BASELINE.json's other configurations (AlexNet, ResNet-18, "
                         "VGG-16 L1 incl. the fc6 variant comparison, ResNet-50 3-bit)")
    ap.add_argument("--no-validate-cpu", dest="validate_cpu", action="store_false",
                    help="skip the one complete real-reference run (AlexNet, about a minute of host time) that checks the "
                         "CPU arm's sampled-and-extrapolated figure")
    ap.add_argument("--cpu-class-seconds", type=float, default=0.3,
                    help="CPU arm: seconds of greedy-loop sampling per layer-shape class and step")
    ap.add_argument("--profile-one-step", action="store_true",
                    help="warm up, then run ONE step between cudaProfilerStart/Stop and exit (for ncu launch lists)")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: route everything else that native libraries may print
    # there (e.g. NCCL's version banner) to stderr, and write the JSON to the real stdout at the end.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    out = sys.stdout
    sys.stdout = os.fdopen(real_stdout, "w")
    try:
        if args.impl == "reference":
            run_reference_arm(args)
        else:
            run_cuda_arm(args)
    finally:
        sys.stdout.flush()


if __name__ == "__main__":
    main()
