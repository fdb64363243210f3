#!/usr/bin/env python
"""AVMoE adapter-layer fwd+bwd tokens/s (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W                 # our CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W  # the reference's CPU path on the host cores

One step = forward + backward of the whole AVE adapter stack (8 adapted blocks x {p1,p2} x {audio,visual} = 32
MoEAdapter calls; Swin-V2-L / HTS-AT token shapes; r=8 G=2 T=32 BN on) on one batch of
synthetic backbone tokens of BT = 16 clips x 10 segments per GPU, bf16 activations, fp32 parameters/router/statistics,
plus the gradient all-reduce when N > 1 (weak scaling: per-GPU batch fixed).  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "avmoe_adapter_fwd_bwd_tokens_per_s"
UNIT = "tokens/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--variant", default="ave")
    ap.add_argument("--clips", type=int, default=16, help="clips per GPU (x10 segments = BT)")
    ap.add_argument("--segments", type=int, default=10)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--f32-gemm", default="tf32", choices=["tf32", "exact"],
                    help="--dtype f32 only: 'tf32' runs the fp32 stream through the same tcgen05 GEMM orchestration as bf16 with TF32 operands "
                         "(fp32 accumulate; the 1e-3 configuration of BASELINE.json config 2), 'exact' keeps the fused kernels whose "
                         "contractions are plain fp32 FMAs (the drop-in's default for fp32 inputs: bit-faithful to torch's fp32 matmul, slow)")
    ap.add_argument("--cpu-sample-clips", type=int, default=2, help="clips in the bounded CPU-baseline sample (2 x 10 = the reference's own batch, AVMOE/AVE/train.sh:32-33)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-eager", action="store_true", help="skip the reported GPU comparator (reference modules in PyTorch eager on the same GPU)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python each step instead of replaying a CUDA graph")
    return ap.parse_args()


def workload_name(args):
    return (f"{args.variant.upper()} adapter stack fwd+bwd, Swin-V2-L/HTS-AT token shapes, batch {args.clips} x {args.segments} "
            f"segments per GPU, {args.dtype}")


# ------------------------------------------------------------------------------------------------ reference (CPU) arm
def cpu_baseline(args, steps, warmup):
    """The reference's own CPU implementation of the path (unmodified reference modules when /root/reference exists,
    else the oracle port) on all host cores, on a bounded sample of the same workload."""
    import torch
    from oracle import ref_loader, stack_oracle
    use_ref = ref_loader.available()
    bt = args.cpu_sample_clips * args.segments
    tps, times, cores = stack_oracle.time_cpu(args.variant, bt, steps, warmup, use_reference=use_ref)
    return {"value": tps, "unit": UNIT, "cores": cores, "kind": "reference" if use_ref else "port",
            "sample": f"{args.variant.upper()} adapter stack fwd+bwd fp32, BT={bt} ({args.cpu_sample_clips} clip x {args.segments} segments), "
                      f"{steps} timed steps after {warmup} warm-up, median {sorted(times)[len(times) // 2]:.3f} s/step"}, times


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base, times = cpu_baseline(args, max(1, args.steps), max(1, min(args.warmup, 1)))
    ms = 1e3 * sorted(times)[len(times) // 2]
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": workload_name(args), "sample": base["sample"]},
            "cpu_baseline": base, "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def gpu_eager_baseline(args, BT):
    """COMPARATOR, reported only: the reference's own adapter modules (staged copies under oracle/_ref or /root/reference; else
    the oracle port) in PyTorch eager on this GPU, same stack / tokens-per-step definition, fp32 (the reference's training
    precision) and under bf16 autocast.  This is what `model.cuda()` gives a reference user."""
    import torch
    from oracle import ref_loader, stack_oracle
    use_ref = ref_loader.available()
    out = {"kind": "reference" if use_ref else "port", "BT": BT, "unit": UNIT}
    for name, ac in (("f32", None), ("bf16_autocast", torch.bfloat16)):
        try:
            tps, ms = stack_oracle.time_gpu_eager(args.variant, BT, 3, 2, use_reference=use_ref, autocast_dtype=ac)
            out[name] = {"value": tps, "ms_per_step": ms}
        except Exception as ex:   # e.g. out of memory for the materialised N x N variants
            out[name] = {"value": None, "error": str(ex)[:200]}
        torch.cuda.empty_cache()
    return out


def stack_roofline(variant, BT, elem_bytes, n_params, tc_peak_tflops, hbm_peak_gbs):
    """SURVEY.md §8(d) stack-level roofline of one step (fwd+bwd) on one GPU:
       algorithmic HBM bytes  = sum over MoEAdapter calls of (5·N·C + 3·Nx·Cx)·s per sample  +  parameters read (4 B) and
                                their gradients written (4 B) once per step;
       resampler FLOPs        = 3 x [2·N·Nx·min(C,Cx) + 2·C·Cx·(Nx if Cx > C else N)] per sample and call (tensor-bound);
       roofline_time          = resampler FLOPs / measured tensor peak  +  algorithmic bytes / measured HBM peak."""
    from avmoe_b200 import stack as S
    nbytes = 0.0
    flops = 0.0
    for stage, blocks in S.VARIANT_SITES[variant]:
        Cv, Nv, Ca, Na = S.STAGE_SHAPES[stage]
        for (C, N, Cx, Nx) in ((Ca, Na, Cv, Nv), (Cv, Nv, Ca, Na)):
            calls = 2 * blocks
            nbytes += calls * BT * (5 * N * C + 3 * Nx * Cx) * elem_bytes
            flops += calls * BT * 3 * (2 * N * Nx * min(C, Cx) + 2 * C * Cx * (Nx if Cx > C else N))
    nbytes += 8.0 * n_params
    t_ms = flops / (tc_peak_tflops * 1e12) * 1e3 + nbytes / (hbm_peak_gbs * 1e9) * 1e3
    return {"algorithmic_bytes_per_step": nbytes, "resampler_flops_per_step": flops, "roofline_time_ms": t_ms,
            "hbm_ms": nbytes / (hbm_peak_gbs * 1e9) * 1e3, "tensor_ms": flops / (tc_peak_tflops * 1e12) * 1e3}


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2]))
                for name, v in zip(names, r[5:9]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from avmoe_b200 import ops, stack as S
    from avmoe_b200.dist import FlatGradientReducer
    from avmoe_b200.optim import FlatParameters, FusedAdam

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    if dtype == torch.float32:      # the drop-in follows torch's switch for fp32 inputs (ops.fp32_on_tensor_cores)
        torch.backends.cuda.matmul.allow_tf32 = args.f32_gemm == "tf32"
    BT = args.clips * args.segments
    model = S.AdapterStack(args.variant, seed=0).to(dev)
    model.train()
    params = [p for p in model.parameters() if p.requires_grad]
    # all adapter + router parameters live in one flat fp32 buffer: the gradient all-reduce runs on its flat gradient twin in
    # place (one multi-tensor gather, no unpack) and the fused Adam consumes it directly with the 1/world folded in
    flat = FlatParameters(params)
    adam = FusedAdam(flat, lr=5e-4)                      # AVMOE/AVE/train.sh: lr 5e-4
    reducer = FlatGradientReducer(flat) if world > 1 else None
    tokens_dev = S.make_stage_tokens(args.variant, BT, dtype, dev, seed=1000 + rank)
    tokens_host = S.make_stage_tokens(args.variant, BT, dtype, None, seed=1000 + rank, pin=True)
    tokens_per_step = S.tokens_per_segment(args.variant) * BT
    h2d_bytes = sum(t.numel() * t.element_size() for pair in tokens_host.values() for t in pair)

    def zero():
        for p in params:
            p.grad = None

    def step(tok, reduce=True):
        zero()
        leaves = {s: (a.detach().requires_grad_(), v.detach().requires_grad_()) for s, (a, v) in tok.items()}
        loss = model(leaves)
        loss.backward()
        if reducer is not None and reduce:
            flat.gather_grads(clear=False)
            reducer.reduce(gather=False)
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- warm-up, then the device-resident timed region ---------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step(tokens_dev)
    # One step = ~2300 small launches; the whole forward+backward is captured once into a CUDA graph and replayed, so the
    # timed region measures the kernels rather than the Python launch path.  (The gradient all-reduce stays outside.)
    graph = graph_loss = None
    launches_per_step = None
    if not args.no_graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                step(tokens_dev)
            torch.cuda.current_stream().wait_stream(side)
            zero()
            graph = torch.cuda.CUDAGraph()
            ops.LAUNCHES = 0
            with torch.cuda.graph(graph):
                leaves = {s: (a.detach().requires_grad_(), v.detach().requires_grad_()) for s, (a, v) in tokens_dev.items()}
                graph_loss = model(leaves)
                graph_loss.backward()
            launches_per_step = ops.LAUNCHES
        except Exception as ex:   # capture is an optimisation of the launch path, never a requirement
            print(f"[bench] CUDA graph capture failed, launching eagerly: {ex}", file=sys.stderr)
            graph = None
            torch.cuda.synchronize()

    def run_step(tok=None):
        """one forward+backward(+all-reduce) over the device-resident tokens (tok: host tensors to copy in first)"""
        if graph is None:
            return step(tokens_dev if tok is None else {s: (a.to(dev, non_blocking=True), v.to(dev, non_blocking=True)) for s, (a, v) in tok.items()})
        if tok is not None:
            for s, (a, v) in tok.items():
                tokens_dev[s][0].copy_(a, non_blocking=True)
                tokens_dev[s][1].copy_(v, non_blocking=True)
        graph.replay()
        if reducer is not None:
            flat.gather_grads(clear=False)
            reducer.reduce(gather=False)
        return graph_loss

    for _ in range(2):
        run_step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ops.LAUNCHES = 0
    ms_total = timed(run_step, args.steps)
    launches = ops.LAUNCHES if graph is None else launches_per_step * args.steps
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = ms_total / args.steps
    value = tokens_per_step * world / (ms_per_step * 1e-3)

    # ---- end-to-end through the public module API with host buffers -----------------------------------------------
    def e2e_step():
        loss = run_step(tokens_host)       # pinned host -> device copies of this step's tokens, then the step
        return float(loss.item())          # device -> host read of the step's result

    if graph is None:
        e2e_step()
        ms_e2e = timed(e2e_step, args.steps) / args.steps
    else:
        # Pipelined: step i+1's tokens stream host -> device (copy stream, staging buffers) while step i computes; a
        # device-to-device copy moves them into the graph's static inputs.  Every step's H2D copy and loss read-back is
        # inside the timed region (K copies for K steps).
        copy_stream = torch.cuda.Stream()
        staging = {s: (torch.empty_like(a), torch.empty_like(v)) for s, (a, v) in tokens_dev.items()}
        landed, consumed = torch.cuda.Event(), torch.cuda.Event()

        def issue_h2d():
            with torch.cuda.stream(copy_stream):
                for s, (a, v) in tokens_host.items():
                    staging[s][0].copy_(a, non_blocking=True)
                    staging[s][1].copy_(v, non_blocking=True)
                landed.record(copy_stream)

        def e2e_loop(n):
            main = torch.cuda.current_stream()
            issue_h2d()
            for i in range(n):
                main.wait_event(landed)
                for s in tokens_dev:
                    tokens_dev[s][0].copy_(staging[s][0], non_blocking=True)
                    tokens_dev[s][1].copy_(staging[s][1], non_blocking=True)
                consumed.record(main)
                if i + 1 < n:
                    copy_stream.wait_event(consumed)
                    issue_h2d()
                loss = run_step()
                float(loss.item())

        e2e_loop(2)
        ms_e2e = timed(lambda: e2e_loop(args.steps), 1) / args.steps
    e2e = {"value": tokens_per_step * world / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes * world,
           "d2h_bytes_per_step": 4 * world, "ms_per_step": ms_e2e}

    # ---- optimizer step on the flat buffers (reported next to the step; the metric itself is fwd+bwd as BASELINE.json defines it)
    flat.gather_grads(clear=False)
    lr_keep = adam.param_groups[0]["lr"]
    adam.param_groups[0]["lr"] = 0.0                     # timing only: the weights (and hence the timed workload) stay as they are
    for _ in range(2):
        adam.step(grad_scale=1.0 / world)
    adam_ms = timed(lambda: adam.step(grad_scale=1.0 / world), 5) / 5
    gather_ms = timed(lambda: flat.gather_grads(clear=False), 5) / 5
    adam.param_groups[0]["lr"] = lr_keep
    optimizer = {"fused_adam_ms": adam_ms, "gather_grads_ms": gather_ms, "params": flat.numel,
                 "adam_GBps": 28.0 * flat.numel / (adam_ms * 1e-3) / 1e9}

    # ---- roofline: per-kernel CUDA-event timing of one extra (untimed) step, on the launching stream ---------------------
    roofline = None
    kernel_table = None
    if rank == 0:
        ops.GEMM_PROFILE, ops.KERNEL_PROFILE = [], []
        step(tokens_dev, reduce=False)     # rank 0 only: no collective in here
        torch.cuda.synchronize()
        gprof, kprof = ops.GEMM_PROFILE, ops.KERNEL_PROFILE
        ops.GEMM_PROFILE = ops.KERNEL_PROFILE = None
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        tc_peak = peaks.get("bf16_tflops_sustained", 1400.0) * (1.0 if dtype == torch.bfloat16 else 0.5)
        peak_src = "MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)"
        agg = {}
        for name, e0, e1, nbytes in kprof:
            a = agg.setdefault(name, dict(ms=0.0, n=0, bytes=0.0))
            a["ms"] += e0.elapsed_time(e1); a["n"] += 1; a["bytes"] += nbytes
        # GEMM launches are split by arithmetic intensity against the machine's ridge point: the resampler GEMMs are
        # tensor-bound, the skinny expert / latent-attention GEMMs stream their big operand once and are HBM-bound.
        ridge = tc_peak * 1e12 / (hbm_peak * 1e9)
        TB, HB = "gemm_tcgen05[tensor-bound]", "gemm_tcgen05[hbm-bound]"
        for e0, e1, fl, nbytes, _tag in gprof:
            a = agg.setdefault(TB if fl / max(nbytes, 1.0) > ridge else HB, dict(ms=0.0, n=0, bytes=0.0, flops=0.0))
            a["ms"] += e0.elapsed_time(e1); a["n"] += 1; a["bytes"] += nbytes; a["flops"] += fl
        if os.environ.get("AVMOE_BENCH_DUMP"):   # diagnostics: per-shape GEMM time of the profiled step
            tags = {}
            for e0, e1, fl, nbytes, tag in gprof:
                t = tags.setdefault(tag, dict(n=0, ms=0.0, GB=0.0, GF=0.0))
                t["n"] += 1; t["ms"] += e0.elapsed_time(e1); t["GB"] += nbytes / 1e9; t["GF"] += fl / 1e9
            with open(os.environ["AVMOE_BENCH_DUMP"], "w") as fh:
                json.dump(dict(sorted(tags.items(), key=lambda kv: -kv[1]["ms"])), fh, indent=1)
        tot_ms = sum(a["ms"] for a in agg.values())
        kernel_table = {k: {"launches": v["n"], "ms": round(v["ms"], 3), "share": round(v["ms"] / tot_ms, 4),
                            **({"GBps": round(v["bytes"] / v["ms"] / 1e6, 1)} if v["bytes"] and k != TB else {}),
                            **({"TFLOPs": round(v["flops"] / v["ms"] / 1e9, 1)} if "flops" in v else {})}
                        for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])}
        dom = max(agg, key=lambda k: agg[k]["ms"])
        d = agg[dom]
        if dom == TB:
            ach = d["flops"] / (d["ms"] * 1e-3) / 1e12
            roofline = {"bound": "tensor", "kernel": "gemm_tcgen05_kernel (token resampler conv_adapter+fc: fwd, dgrad, wgrad)", "achieved": ach,
                        "peak": tc_peak, "unit": "TFLOP/s", "frac": ach / tc_peak, "traffic": None}
        else:
            ach = d["bytes"] / (d["ms"] * 1e-3) / 1e9
            name = "gemm_tcgen05_kernel (skinny expert-projection / latent-attention GEMMs)" if dom == HB else f"{dom}_kernel"
            traffic = traffic_src = None
            if dom == HB:   # DRAM bytes per GEMM launch from the ncu pass over one step of the same command (profiles/r02_dram_step.json:
                            # all tcgen05 GEMM launches of the step, tensor-bound ones included), else round 1's measured dram/algorithmic ratio
                try:
                    dr = json.load(open(os.path.join(ROOT, "profiles", "r02_dram_step.json")))
                    gk = [v for k, v in dr["by_kernel"].items() if "gemm_tcgen05" in k][0]
                    traffic = gk["GB"] * 1e9 / gk["n"]
                    traffic_src = ("profiles/r02_dram_step.json: ncu dram__bytes_read+write of the %d tcgen05 GEMM launches of one step / launches "
                                   "(all GEMM launches; round-1 capture of the skinny group alone: 0.94 x algorithmic)" % gk["n"])
                except Exception:
                    try:
                        tr = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic_gemm.json")))
                        traffic = tr["dram_to_algorithmic_ratio"] * d["bytes"] / max(d["n"], 1)
                        traffic_src = "profiles/r01_traffic_gemm.json: measured dram/algorithmic ratio x algorithmic bytes per launch"
                    except Exception:
                        pass
            roofline = {"bound": "hbm", "kernel": name, "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                        "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes_per_launch": d["bytes"] / max(d["n"], 1)}
        gt = agg.get(TB)
        st = stack_roofline(args.variant, BT, 2 if dtype == torch.bfloat16 else 4, sum(p.numel() for p in params), tc_peak, hbm_peak)
        st["achieved"] = st["roofline_time_ms"] / ms_per_step          # fraction of the stack-level roofline (north-star bar: 0.60)
        st["kernel_moved_bytes_per_step"] = sum(a["bytes"] for a in agg.values())   # bytes the kernels' own operands add up to
        st["moved_over_algorithmic"] = st["kernel_moved_bytes_per_step"] / st["algorithmic_bytes_per_step"]
        try:    # whole-step DRAM traffic from an ncu pass of the same command (tools/dram_step.py), when one is committed
            dr = json.load(open(os.path.join(ROOT, "profiles", "r02_dram_step.json")))
            if dr.get("variant") == args.variant and dr.get("BT") == BT:
                st["dram_bytes_per_step_ncu"] = dr["dram_bytes"]
                st["dram_over_algorithmic"] = dr["dram_bytes"] / st["algorithmic_bytes_per_step"]
                st["dram_source"] = "profiles/r02_dram_step.json"
        except Exception:
            pass
        roofline["stack"] = st
        roofline.update({"launches_per_step": d["n"], "avg_launch_ms": d["ms"] / max(d["n"], 1), "share_of_kernel_time": d["ms"] / tot_ms,
                         "peak_source": peak_src, "ridge_flop_per_byte": ridge,
                         "gemm_tensor_bound": None if not gt else {
                             "achieved_TFLOPs": gt["flops"] / (gt["ms"] * 1e-3) / 1e12, "peak_TFLOPs": tc_peak,
                             "frac": gt["flops"] / (gt["ms"] * 1e-3) / 1e12 / tc_peak, "launches_per_step": gt["n"], "ms_per_step": gt["ms"],
                             "algorithmic_flops_per_step": gt["flops"]}})

    # ---- CPU baseline (rank 0, N=1 only), bounded sample --------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu, _ = cpu_baseline(args, 2, 1)
        except Exception as ex:  # the baseline is a report, never a dependency of the product path
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {ex}"}

    launch_mode = "cuda-graph replay" if graph is not None else "eager python launches"
    gpu_eager = None
    if rank == 0 and world == 1 and not args.no_gpu_eager:
        graph = graph_loss = None
        model.zero_grad(set_to_none=True)
        torch.cuda.empty_cache()
        try:
            gpu_eager = gpu_eager_baseline(args, BT)
        except Exception as ex:
            gpu_eager = {"error": str(ex)[:200]}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16" if dtype == torch.bfloat16 else "f32", "data": "synthetic",
                "config": {"workload": workload_name(args), "variant": args.variant,
                           **({"f32_gemm": args.f32_gemm} if dtype == torch.float32 else {}), "BT_per_gpu": BT, "tokens_per_step_per_gpu": tokens_per_step,
                           "adapter_calls_per_step": 4 * len(model.sites), "launch": launch_mode, "parallelism": f"dp{world} (batch-sharded, grad all-reduce)" if world > 1 else "single GPU",
                           "l2": f"inputs+weights per step ({(h2d_bytes + sum(p.numel() * 4 for p in params)) / 1e6:.0f} MB) exceed the 126 MB L2; no explicit flush"},
                "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "optimizer": optimizer, "kernels": kernel_table}
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if gpu_eager is not None:
            line["gpu_eager_baseline"] = gpu_eager
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
