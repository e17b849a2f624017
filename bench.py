#!/usr/bin/env python
"""Benchmark of the hot path: integrand-evaluations/s of the fused Clenshaw-Curtis kernel.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg4]

One "step" = one pass of the hot path over one synthetic batch: for every (sample, dimension) slot
the integral over the Q+1 CC nodes plus the Jacobian point f(x, h)  (BASELINE.json north_star).
Metric: integrand-evals/s with evals := B*D*Q (BASELINE.json's convention; the kernel really
evaluates B*D*(Q+2) rows, reported in config.rows_per_step).

Our arm prints ONE JSON line with value (inputs resident in HBM), e2e (pinned host buffers in, results
back on the host, through the public API), roofline (tensor-pipe bound: algorithmic FLOP / CUDA-event
time / measured sustained bf16 peak) and cpu_baseline (the C oracle port on the host cores, bounded
sample).  `--impl reference` times that CPU port alone with all host threads.
Multi-GPU: one process per GPU under torchrun, batch-sharded, no data-path collective ("weak").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (B per GPU, D, E, hidden, Q, layout)
    "cfg1": dict(B=100, D=1, E=2, hidden=[64, 64, 64], Q=50, layout="contig",
                 label="MonotonicMLP smoke B=100 Q=50 [64,64,64]"),
    "cfg2": dict(B=10000, D=2, E=10, hidden=[100, 100, 100, 100], Q=50, layout="strided",
                 label="Toy 2-moons UMNN-MAF D=2 Q=50 B=10000"),
    "cfg3": dict(B=10000, D=6, E=30, hidden=[200, 200, 200], Q=50, layout="strided",
                 label="UCI POWER D=6 [200,200,200] Q=50 B=10000"),
    "cfg4": dict(B=65536, D=63, E=30, hidden=[200, 200, 200], Q=100, layout="strided",
                 label="BSDS300-shape synthetic D=63 [200,200,200] Q=100 B=65536"),
    "cfg5": dict(B=100, D=784, E=30, hidden=[100, 50, 50, 50, 50], Q=50, layout="strided",
                 label="MNIST UMNN-MAF D=784 [100,50,50,50,50] Q=50 B=100"),
}


def flop_per_eval(widths):
    return 2 * sum(a * b for a, b in zip(widths[:-1], widths[1:]))


def load_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm_gbs=p["hbm_gbs"], bf16_tflops=p["bf16_tflops"],
                    bf16_tflops_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi sampler running during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu_index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, power, reasons = [], [], [], set()
        for ln in self.lines:
            f = [c.strip() for c in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(power)),
                "samples": len(sm), "reasons": sorted(reasons)}


def make_problem(cfg, seed):
    from oracle import umnn_oracle as orc
    spec = orc.MLPSpec(tuple([1 + cfg["E"]] + cfg["hidden"] + [1]),
                       orc.HIDDEN_LEAKY if cfg["layout"] == "strided" else orc.HIDDEN_RELU, orc.OUT_ELU_PLUS_1)
    flat = orc.synth_params(spec, seed)
    return spec, flat


def cpu_arm(cfg, spec, flat, budget_s=12.0, n_threads=0):
    """The C port of the reference path on the host cores over a bounded sample of the workload."""
    from oracle import c_binding, umnn_oracle as orc
    D, E, Q = cfg["D"], cfg["E"], cfg["Q"]
    Hh = E * D if cfg["layout"] == "strided" else E
    threads = n_threads or c_binding.max_threads()

    def run(Bs):
        x0, x, h, _ = orc.synth_inputs(Bs, D, Hh, 1)
        t0 = time.perf_counter()
        c_binding.cc_forward(spec, flat, x0, x, h, Q, cfg["layout"], want_f=True, n_threads=threads)
        return time.perf_counter() - t0

    probe_B = max(1, min(cfg["B"], max(threads, 16384 // max(1, D * (Q + 2) // 64))))
    run(min(probe_B, 8))                      # warm-up (thread creation, page faults)
    t_probe = run(probe_B)
    Bs = int(min(cfg["B"], max(probe_B, probe_B * budget_s / max(t_probe, 1e-4))))
    t = run(Bs)
    evals = Bs * D * Q
    return {"value": evals / t, "unit": "integrand-evals/s", "cores": threads, "kind": "port",
            "sample": f"oracle/umnn_oracle.c (pthreads x{threads}) on B={Bs} of {cfg['B']} samples of the same "
                      f"workload, {t:.2f} s"}, t, Bs


def run_reference(args, cfg, name):
    """--impl reference: the CPU port with all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    spec, flat = make_problem(cfg, 0)
    from oracle import c_binding, umnn_oracle as orc
    threads = c_binding.max_threads()
    D, E, Q = cfg["D"], cfg["E"], cfg["Q"]
    Hh = E * D if cfg["layout"] == "strided" else E
    # size one step at ~2 s of CPU work so K + W steps end within a couple of minutes
    base, t_base, Bs = cpu_arm(cfg, spec, flat, budget_s=2.0)
    x0, x, h, _ = orc.synth_inputs(Bs, D, Hh, 1)
    for _ in range(args.warmup):
        c_binding.cc_forward(spec, flat, x0, x, h, Q, cfg["layout"], want_f=True, n_threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        c_binding.cc_forward(spec, flat, x0, x, h, Q, cfg["layout"], want_f=True, n_threads=threads)
    dt = time.perf_counter() - t0
    value = args.steps * Bs * D * Q / dt
    sample = (f"oracle/umnn_oracle.c (C restatement of the reference's CPU path, pthreads x{threads}); each step = "
              f"B={Bs} of {cfg['B']} samples of the workload")
    out = {"impl": "reference", "metric": "integrand-evals/sec (B*D*Q)", "value": value, "unit": "integrand-evals/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": f"{name}: {cfg['label']}", "sample_batch": Bs, "D": D, "Q": Q,
                      "hidden": cfg["hidden"], "E": E},
           "cpu_baseline": {"value": value, "unit": "integrand-evals/s", "cores": threads, "kind": "port",
                            "sample": sample},
           "e2e": {"value": value, "unit": "integrand-evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


AUTO_TC_DEFAULT = "fp16x3"   # what UMNN_PREC_AUTO resolves to on the tensor cores (umnn_abi.cu: auto_tc_precision)


def torch_route_probe(net, cfg, dev, Bs=256, reps=3):
    """The reference-equivalent PyTorch-CUDA formulation (our torch route = the ParallelNeuralIntegral
    algorithm as torch ops: expand / cat / transpose / Linear / activations / weighted sum) on a sub-batch."""
    import torch
    from umnn_b200.integral import integrate
    D, E, Q = cfg["D"], cfg["E"], cfg["Q"]
    Hh = E * D if cfg["layout"] == "strided" else E
    g = torch.Generator(device=dev).manual_seed(7)
    x = 2 * torch.randn(Bs, D, device=dev, generator=g)
    h = torch.randn(Bs, Hh, device=dev, generator=g)
    x0 = torch.zeros_like(x)
    with torch.no_grad():
        for _ in range(2):
            integrate(x0, Q, (x - x0) / Q, net, h)
            net(x, h)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            integrate(x0, Q, (x - x0) / Q, net, h)
            net(x, h)
        e.record()
        torch.cuda.synchronize()
    ms = s.elapsed_time(e) / reps
    return {"value": Bs * D * Q / (ms * 1e-3), "unit": "integrand-evals/s", "sample_batch": Bs,
            "what": "reference algorithm as PyTorch-CUDA ops (fp32, TF32 off), sub-batch, same B200"}


def backward_probe(net, cfg, dev, reps=3):
    """Informational: the Leibniz backward (d_x0, d_x, d_h, d_params) on a sub-batch through the three paths."""
    import torch
    from umnn_b200 import _native, kernel
    from umnn_b200.integral import _integrate_grads_chunked
    D, E, Q = cfg["D"], cfg["E"], cfg["Q"]
    Hh = E * D if cfg["layout"] == "strided" else E
    Bs = max(1, min(cfg["B"], 2_000_000 // (D * (Q + 3))))
    g = torch.Generator(device=dev).manual_seed(11)
    x = 2 * torch.randn(Bs, D, device=dev, generator=g)
    h = torch.randn(Bs, Hh, device=dev, generator=g)
    go = torch.randn(Bs, D, device=dev, generator=g)
    x0 = torch.zeros_like(x)
    ks = net.kernel_spec()
    res = {"sample_batch": Bs, "rows": Bs * D * (Q + 3)}

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / reps

    for label, prec in (("tensor_core_default_ms", _native.PREC_AUTO), ("bf16x3_tensor_core_ms", _native.PREC_BF16X3),
                        ("fp32_ffma_ms", _native.PREC_FP32)):
        if _native.lib().umnn_workspace_bytes(kernel.make_desc(ks, x, Q, prec), 1) > 0:
            res[label] = timed(lambda: kernel.cc_backward(ks, x0, x, h, go, Q, precision=prec))
    res["torch_ops_reference_algorithm_ms"] = timed(lambda: _integrate_grads_chunked(x0, x, net, h, Q, go, False))
    return res


def run_ours(args, cfg, name):
    import torch
    import torch.distributed as dist
    from oracle import umnn_oracle as orc
    from umnn_b200 import IntegrandNN, IntegrandNetwork, _native, cc_integrate, cc_integrate_host

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    _native.lib()  # the CUDA extension must be there: no fallback

    spec, flat = make_problem(cfg, 0)
    B, D, E, Q = cfg["B"], cfg["D"], cfg["E"], cfg["Q"]
    Hh = E * D if cfg["layout"] == "strided" else E
    # which kernel family will serve this shape (the library resolves `auto` the same way)
    probe = _native.make_desc(_native.LAYOUT_STRIDED_D if cfg["layout"] == "strided" else _native.LAYOUT_CONTIG, B, D, E,
                              spec.widths, _native.ACT_LEAKY_RELU if cfg["layout"] == "strided" else _native.ACT_RELU,
                              _native.OUT_ELU_PLUS_1, Q, _native.PREC_BF16X3)
    tc_ok = _native.lib().umnn_packed_params_bytes(probe) > 0
    use_tc = args.precision in ("bf16x3", "fp16x3") or (args.precision == "auto" and tc_ok)
    # which operand split the tensor-core kernel uses (the library resolves `auto` through UMNN_B200_AUTO_TC)
    auto_tc = os.environ.get("UMNN_B200_AUTO_TC", AUTO_TC_DEFAULT).lower()
    split = args.precision if args.precision != "auto" else auto_tc
    pad = lambda w: (w + 2 + 15) // 16 * 16
    issued_per_row = 3 * 2 * sum(pad(a) * pad(b) for a, b in zip(spec.widths[1:-2], spec.widths[2:-1])) if use_tc else 0
    net = IntegrandNetwork(D, 1 + E, cfg["hidden"], 1) if cfg["layout"] == "strided" else IntegrandNN(1 + E, cfg["hidden"])
    off = 0
    with torch.no_grad():
        for p in net.parameters():
            p.copy_(torch.from_numpy(flat[off:off + p.numel()].copy()).view_as(p))
            off += p.numel()
    net.to(dev).eval()     # inference-style step: packed parameters are cached, ONE kernel launch per step

    # every rank owns its own shard of the global batch (B samples per GPU): weak scaling, no collective
    gen = torch.Generator(device="cpu").manual_seed(1000 + rank)
    x_host = (2 * torch.randn(B, D, generator=gen)).pin_memory()
    h_host = torch.randn(B, Hh, generator=gen).pin_memory()
    out_host = torch.empty(B, D).pin_memory()
    fx_host = torch.empty(B, D).pin_memory()
    x = x_host.to(dev)
    h = h_host.to(dev)

    def step_resident():
        return cc_integrate(net, None, x, h, Q, want_fx=True)

    def step_e2e():
        # public host-buffer entry: pinned inputs up, fused launch, results down (chunked so the copies overlap)
        cc_integrate_host(net, x_host, h_host, Q, want_fx=True, out=out_host, fx_out=fx_host)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            fn()
        e.record()
        barrier()
        ms = s.elapsed_time(e)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_total = timed(step_resident, args.steps)
    clocks = sampler.stop() if sampler else None

    for _ in range(2):
        step_e2e()
    e2e_steps = max(2, min(args.steps, 5))
    ms_e2e = timed(step_e2e, e2e_steps)

    # parity of what was just timed: a few samples of this rank's shard against the C oracle
    o, f, _ = step_resident()
    torch.cuda.synchronize()
    from oracle import c_binding
    n_chk = min(B, 4 if D > 16 else 32)
    ref, rfx, _ = c_binding.cc_forward(spec, flat, np.zeros((n_chk, D), np.float32), x_host[:n_chk].numpy(),
                                       h_host[:n_chk].numpy(), Q, cfg["layout"])
    got = o[:n_chk].cpu().numpy()
    rel = float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-6)))
    jac_abs = float(np.max(np.abs(np.log(f[:n_chk].cpu().numpy() + 1e-10) - np.log(rfx + 1e-10))))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    evals_per_step = B * D * Q * world
    rows_per_step = B * D * (Q + 2) * world
    ms_per_step = ms_total / args.steps
    value = evals_per_step / (ms_per_step * 1e-3)
    e2e_value = evals_per_step / (ms_e2e / e2e_steps * 1e-3)
    peaks = load_peaks()
    fpe = flop_per_eval(spec.widths)
    # per-GPU kernel: rows of one launch x FLOP per row / average launch duration (one launch per step)
    achieved_tflops = (rows_per_step / world) * fpe / (ms_per_step * 1e-3) / 1e12
    peak = peaks["bf16_tflops_sustained"]
    slot_bytes = 4 * (E + 1) + 8            # x, h read (x0 = NULL); integral, f(x) written
    traffic = None
    tpath = os.path.join(REPO, "profiles", "traffic.json")
    if os.path.exists(tpath):
        # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch from an `ncu --set full` capture of this
        # workload / kernel (scripts/gpu_round.sh), recorded per slot so it scales with the batch
        rec = json.load(open(tpath)).get(f"{name}:{'tc' if use_tc else 'fp32'}")
        if rec:
            traffic = rec["dram_bytes_per_slot"] * B * D
    out = {
        "metric": "integrand-evals/sec (B*D*Q)", "value": value, "unit": "integrand-evals/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": f"{split} (16-bit hi+lo operand split, fp32 accumulate)" if use_tc else "f32",
        "data": "synthetic",
        "config": {"workload": f"{name}: {cfg['label']}", "per_gpu_batch": B, "global_batch": B * world, "D": D,
                   "E": E, "Q": Q, "hidden": cfg["hidden"], "parallelism": f"batch-shard x{world}, no collective",
                   "rows_per_step": rows_per_step, "precision": args.precision,
                   "kernel": "cc_forward_tc (tcgen05 cta_group::2)" if use_tc else "cc_forward_fp32 (FFMA)",
                   "l2": f"inputs {(x.numel() + h.numel()) * 4 / 1e6:.0f} MB per step "
                         + ("> 126 MB L2 (no flush needed)" if (x.numel() + h.numel()) * 4 > 126e6 else "< L2: resident")},
        "e2e": {"value": e2e_value, "unit": "integrand-evals/s", "ms_per_step": ms_e2e / e2e_steps,
                "h2d_bytes_per_step": int((x_host.numel() + h_host.numel()) * 4),
                "d2h_bytes_per_step": int((out_host.numel() + fx_host.numel()) * 4),
                "api": "umnn_b200.cc_integrate_host on pinned host tensors (H2D + fused kernel + D2H every step, "
                       "pipelined over batch chunks)"},
        # fp16x3: the fused kernel + its guarded bf16 re-run (a no-op launch unless an activation left the fp16 range)
        "gpu_launches": args.steps * world * (2 if (use_tc and split == "fp16x3") else 1),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "achieved": achieved_tflops, "peak": peak, "unit": "TFLOP/s",
                     "frac": achieved_tflops / peak, "traffic": traffic,
                     "algorithmic_bytes": B * D * slot_bytes,
                     "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peaks['source']})",
                     "flop_per_row": fpe, "rows_per_launch": rows_per_step // world,
                     "hbm_sanity_gbs": (B * D * slot_bytes) / (ms_per_step * 1e-3) / 1e9,
                     "issued_tensor_tflops": (rows_per_step / world) * issued_per_row / (ms_per_step * 1e-3) / 1e12,
                     "issued_tensor_frac": (rows_per_step / world) * issued_per_row / (ms_per_step * 1e-3) / 1e12 / peak,
                     "note": "achieved = algorithmic fp32 FLOP (2*sum in*out per row) / time; the hi+lo operand split issues 3 "
                             "padded 16-bit MMAs per algorithmic MAC on the hidden layers, so frac is capped near 1/3.3; "
                             "issued_tensor_* counts the 16-bit FLOP actually sent to the tensor pipe"},
        "parity": {"integral_max_rel_err_vs_oracle": rel, "log_jac_max_abs_err_vs_oracle": jac_abs, "samples": n_chk},
    }
    if world == 1 and not args.no_cpu:
        out["cpu_baseline"] = cpu_arm(cfg, spec, flat)[0]
        out["aux"] = {"torch_cuda_reference_algorithm": torch_route_probe(net, cfg, dev),
                      "backward": backward_probe(net, cfg, dev)}
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch (debugging only)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--precision", default=os.environ.get("UMNN_B200_PRECISION", "auto"), choices=["auto", "fp32", "bf16x3", "fp16x3"],
                    help="kernel family: auto = tensor cores (fp16x3 split, guarded bf16x3 re-run) when the shape fits, else FP32 FFMA")
    args = ap.parse_args()
    os.environ["UMNN_B200_PRECISION"] = args.precision
    cfg = dict(WORKLOADS[args.workload])
    if args.batch:
        cfg["B"] = args.batch
    if args.impl == "reference":
        run_reference(args, cfg, args.workload)
    else:
        run_ours(args, cfg, args.workload)


if __name__ == "__main__":
    main()
