#!/usr/bin/env python
"""Benchmark of the hot path: integrand-evaluations/s of the fused Clenshaw-Curtis kernel.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg4]

One "step" = one pass of the hot path over one synthetic batch: for every (sample, dimension) slot
the integral over the Q+1 CC nodes plus the Jacobian point f(x, h)  (BASELINE.json north_star).
Metric: integrand-evals/s with evals := B*D*Q (BASELINE.json's convention; the kernel really
evaluates B*D*(Q+2) rows, reported in config.rows_per_step).

Our arm prints ONE JSON line with
  value         inputs resident in HBM, CUDA events, max over ranks
  e2e           pinned host buffers in, results back on the host, through the public API, over the same K steps
  roofline      tensor-pipe bound: algorithmic FLOP / CUDA-event time / measured sustained bf16 peak
  cpu_baseline  the UNMODIFIED reference (baseline/_ref/UMNN, scripts/vendor_reference.py) on the host cores, bounded
                sample (N=1 only); aux.cpu_port is the C restatement of the same path (oracle/umnn_oracle.c)
  aux           reference_torch_cuda: the unmodified reference's ParallelNeuralIntegral on THIS GPU (fp32, TF32 off,
                batch-chunked) -- the denominator of the north star's ">= 10x"; backward probe; train_step legs
                (UMNNMAFFlow.compute_ll forward + backward + bucketed NCCL all-reduce overlapped with the backward
                + Adam, configs 3 and 5) with the exposed all-reduce time.
`--impl reference` times the unmodified reference's CPU path alone with all host threads.

Multi-GPU: one process per GPU under torchrun.  STRONG scaling: the BASELINE batch (config 4: 65536 samples) is
sharded over the ranks (65536/N per GPU), no data-path collective in the forward; the training legs add the
gradient all-reduce.
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
    # name: (GLOBAL batch B, D, E, hidden, Q, layout)
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

# training legs: UMNNMAFFlow shaped like the reference's drivers (UCIExperiments.py:205-215, MNISTExperiment.py:239-245)
TRAIN_FLOWS = {
    "cfg3": dict(D=6, E=30, Q=50, hid=[200, 200, 200], made=[512, 512], blocks=5, B=10000,
                 label="UCI POWER flow: 5 blocks, MADE [512,512], integrand [200]^3, global batch 10000"),
    "cfg5": dict(D=784, E=30, Q=50, hid=[100, 50, 50, 50, 50], made=[1024, 1024, 1024], blocks=5, B=100,
                 label="MNIST flow: 5 blocks, MADE [1024]^3, integrand [100,50,50,50,50], global batch 100"),
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


# --------------------------------------------------------------------------------------------------------------------
# the unmodified reference (baseline/_ref/UMNN): loaded under its own package name, never mixed with umnn_b200
# --------------------------------------------------------------------------------------------------------------------
def load_reference():
    ref_dir = os.path.join(REPO, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "UMNN")):
        return None      # vendored by __graft_entry__.build() / scripts/vendor_reference.py in the build container
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    import UMNN  # noqa: F401  (the reference package as its wheel would install it)
    return UMNN


def reference_integrand(ref, cfg, flat, device):
    """The reference's own integrand module with the benchmark's weights (same nn.Sequential parameter order)."""
    import torch
    if cfg["layout"] == "strided":
        from UMNN.UMNNMAF import IntegrandNetwork as RefNet
        net = RefNet(cfg["D"], 1 + cfg["E"], list(cfg["hidden"]), 1, act_func="ELU", device=device)
    else:
        net = ref.IntegrandNN(1 + cfg["E"], list(cfg["hidden"]))
    off = 0
    with torch.no_grad():
        for p in net.parameters():
            p.copy_(torch.from_numpy(flat[off:off + p.numel()].copy()).view_as(p))
            off += p.numel()
    return net.to(device)


def reference_step(ref, net, x, h, Q):
    """What one bench step is in the reference: ParallelNeuralIntegral.apply (ParallelNeuralIntegral.py:97-108) from
    x0 = 0 as in UMNNMAF.forward (UMNNMAF.py:77,115), plus the Jacobian point of compute_log_jac (:136-139)."""
    import torch
    with torch.no_grad():
        x0 = torch.zeros_like(x)
        z = ref.ParallelNeuralIntegral.apply(x0, x, net, None, h, Q)
        jac = net(x, h)
    return z, jac


def host_threads():
    """All the host cores this process may use.  torchrun exports OMP_NUM_THREADS=1 to its workers, which would
    silently turn the CPU arms into single-thread runs: the arms set the thread count explicitly instead."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def reference_cpu_arm(cfg, flat, budget_s, n_threads=0):
    """The unmodified reference on the host cores, bounded sample of the workload, batch-chunked like the reference's
    own drivers sub-batch (MNISTExperiment.py:46,160).  Returns (evals/s, info) or None if the reference is absent."""
    import torch
    ref = load_reference()
    if ref is None:
        return None
    torch.set_num_threads(n_threads or host_threads())
    threads = torch.get_num_threads()
    D, E, Q = cfg["D"], cfg["E"], cfg["Q"]
    Hh = E * D if cfg["layout"] == "strided" else E
    net = reference_integrand(ref, cfg, flat, "cpu")
    rows_per_sample = D * (Q + 1)
    chunk = max(1, min(cfg["B"], (1 << 21) // max(1, rows_per_sample * max(cfg["hidden"]) // 64)))   # ~0.5 GB of activations
    g = torch.Generator().manual_seed(7)

    def run(n_chunks):
        t0 = time.perf_counter()
        for _ in range(n_chunks):
            x = 2 * torch.randn(chunk, D, generator=g)
            h = torch.randn(chunk, Hh, generator=g)
            reference_step(ref, net, x, h, Q)
        return time.perf_counter() - t0

    run(1)
    t1 = run(1)
    n_chunks = int(max(1, min((cfg["B"] + chunk - 1) // chunk, budget_s / max(t1, 1e-4))))
    t = run(n_chunks)
    Bs = n_chunks * chunk
    value = Bs * D * Q / t
    return value, {"cores": threads, "sample_batch": Bs, "chunk": chunk, "seconds": t,
                   "sample": f"unmodified reference ParallelNeuralIntegral.apply + IntegrandNetwork (baseline/_ref/UMNN, PyTorch "
                             f"CPU, {threads} threads) on B={Bs} of {cfg['B']} samples of the same workload in chunks of "
                             f"{chunk}, {t:.2f} s"}


def cpu_port_arm(cfg, spec, flat, budget_s=6.0, n_threads=0):
    """The C restatement of the reference path on the host cores over a bounded sample (second CPU figure)."""
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
    run(min(probe_B, 8))
    t_probe = run(probe_B)
    Bs = int(min(cfg["B"], max(probe_B, probe_B * budget_s / max(t_probe, 1e-4))))
    t = run(Bs)
    return {"value": Bs * D * Q / t, "unit": "integrand-evals/s", "cores": threads, "kind": "port",
            "sample": f"oracle/umnn_oracle.c (pthreads x{threads}) on B={Bs} of {cfg['B']} samples, {t:.2f} s"}


def run_reference(args, cfg, name):
    """--impl reference: the reference's own CPU implementation of the path with all host threads; each step is a
    bounded sample of the workload.  Falls back to the C port (kind "port") only if baseline/_ref is absent."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    spec, flat = make_problem(cfg, 0)
    D, E, Q = cfg["D"], cfg["E"], cfg["Q"]
    Hh = E * D if cfg["layout"] == "strided" else E
    ref = load_reference()
    if ref is not None:
        torch.set_num_threads(host_threads())
        threads = torch.get_num_threads()
        net = reference_integrand(ref, cfg, flat, "cpu")
        # size one step at ~2 s of CPU work so K + W steps end within a couple of minutes
        v0, info = reference_cpu_arm(cfg, flat, budget_s=2.0)
        Bs, chunk = info["sample_batch"], info["chunk"]
        g = torch.Generator().manual_seed(1)
        xs = [2 * torch.randn(chunk, D, generator=g) for _ in range(Bs // chunk)]
        hs = [torch.randn(chunk, Hh, generator=g) for _ in range(Bs // chunk)]

        def step():
            for x, h in zip(xs, hs):
                reference_step(ref, net, x, h, Q)
        kind = "reference"
        sample = (f"unmodified reference (baseline/_ref/UMNN: ParallelNeuralIntegral.apply + IntegrandNetwork.forward, PyTorch CPU, "
                  f"{threads} threads); each step = B={Bs} of {cfg['B']} samples of the workload in chunks of {chunk}")
    else:
        from oracle import c_binding, umnn_oracle as orc
        threads = c_binding.max_threads()
        base = cpu_port_arm(cfg, spec, flat, budget_s=2.0)
        Bs = int(base["sample"].split("B=")[1].split(" ")[0])
        x0, x, h, _ = orc.synth_inputs(Bs, D, Hh, 1)

        def step():
            c_binding.cc_forward(spec, flat, x0, x, h, Q, cfg["layout"], want_f=True, n_threads=threads)
        kind = "port"
        sample = f"baseline/_ref absent: oracle/umnn_oracle.c (pthreads x{threads}); each step = B={Bs} of {cfg['B']} samples"
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = args.steps * Bs * D * Q / dt
    out = {"impl": "reference", "metric": "integrand-evals/sec (B*D*Q)", "value": value, "unit": "integrand-evals/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": f"{name}: {cfg['label']}", "sample_batch": Bs, "D": D, "Q": Q,
                      "hidden": cfg["hidden"], "E": E},
           "cpu_baseline": {"value": value, "unit": "integrand-evals/s", "cores": threads, "kind": kind, "sample": sample},
           "e2e": {"value": value, "unit": "integrand-evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


# --------------------------------------------------------------------------------------------------------------------
# probes of our arm (N = 1)
# --------------------------------------------------------------------------------------------------------------------
def reference_cuda_probe(cfg, flat, dev, reps=3):
    """The UNMODIFIED reference on this GPU: ParallelNeuralIntegral.apply + IntegrandNetwork.forward from
    baseline/_ref/UMNN, fp32, TF32 off (PyTorch default matmul precision), batch-chunked because the un-chunked call
    materialises ~33 MB per sample at config 4.  The north star's ">= 10x" denominator."""
    import torch
    ref = load_reference()
    if ref is None:
        return {"unavailable": "baseline/_ref/UMNN not present (scripts/vendor_reference.py needs /root/reference)"}
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    D, E, Q = cfg["D"], cfg["E"], cfg["Q"]
    Hh = E * D if cfg["layout"] == "strided" else E
    net = reference_integrand(ref, cfg, flat, dev)
    per_sample = 4.0 * D * (Q + 1) * (2 * (1 + E) + 3 * max(cfg["hidden"]))      # bytes of live intermediates, roughly
    free = torch.cuda.mem_get_info(dev)[0]
    chunk = int(max(1, min(cfg["B"], 2048, 0.25 * free / per_sample)))
    g = torch.Generator(device=dev).manual_seed(7)
    x = 2 * torch.randn(chunk, D, device=dev, generator=g)
    h = torch.randn(chunk, Hh, device=dev, generator=g)
    for _ in range(2):
        reference_step(ref, net, x, h, Q)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        reference_step(ref, net, x, h, Q)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / reps
    peak = torch.cuda.max_memory_allocated(dev) / 2**30
    del net, x, h
    torch.cuda.empty_cache()
    return {"value": chunk * D * Q / (ms * 1e-3), "unit": "integrand-evals/s", "chunk_batch": chunk, "ms_per_chunk": ms,
            "peak_GiB": round(peak, 2),
            "what": "unmodified reference (baseline/_ref/UMNN) ParallelNeuralIntegral.apply + IntegrandNetwork.forward as "
                    "PyTorch-CUDA ops on this GPU, fp32, TF32 off, one chunk of the workload timed and scaled linearly"}


def backward_probe(net, cfg, dev, reps=3):
    """Informational: the Leibniz backward (d_x0, d_x, d_h, d_params) on a sub-batch through the native paths."""
    import torch
    from umnn_b200 import _native, kernel
    from umnn_b200.integral import _integrate_grads_chunked
    D, E, Q = cfg["D"], cfg["E"], cfg["Q"]
    Hh = E * D if cfg["layout"] == "strided" else E
    Bs = max(1, min(cfg["B"], 4_000_000 // (D * (Q + 3))))
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

    for label, prec in (("tensor_core_default_ms", _native.PREC_AUTO), ("fp32_ffma_ms", _native.PREC_FP32)):
        if _native.lib().umnn_workspace_bytes(kernel.make_desc(ks, x, Q, prec), 1) > 0:
            res[label] = timed(lambda: kernel.cc_backward(ks, x0, x, h, go, Q, precision=prec))
    if "tensor_core_default_ms" in res:
        fpe = flop_per_eval([1 + E] + cfg["hidden"] + [1])
        res["tensor_core_algorithmic_tflops"] = 3 * fpe * res["rows"] / (res["tensor_core_default_ms"] * 1e-3) / 1e12
    n_small = max(1, Bs // 8)
    res["torch_ops_reference_algorithm_ms_scaled"] = 8 * timed(
        lambda: _integrate_grads_chunked(x0[:n_small], x[:n_small], net, h[:n_small], Q, go[:n_small], False))
    return res


def train_step_leg(tcfg, dev, world, rank, steps, warmup):
    """UMNNMAFFlow.compute_ll forward + backward + gradient all-reduce (NCCL, bucketed, launched from the gradient
    hooks so it overlaps the rest of the backward) + Adam -- the data-parallel form of UCIExperiments.py:129-144.
    Strong scaling: the global batch is sharded.  Returns ms/step (max over ranks) and the exposed all-reduce ms."""
    import torch
    import torch.distributed as dist
    from umnn_b200 import UMNNMAFFlow
    from umnn_b200.distributed import BucketedGradientAllReduce, shard_bounds
    torch.manual_seed(0)                                    # same initial weights on every rank
    model = UMNNMAFFlow(nb_flow=tcfg["blocks"], nb_in=tcfg["D"], hidden_derivative=tcfg["hid"],
                        hidden_embedding=tcfg["made"], embedding_s=tcfg["E"], nb_steps=tcfg["Q"], solver="CCParallel",
                        device=dev).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    n_params = sum(p.numel() for p in model.parameters() if p.requires_grad)
    bucket = max(1 << 20, min(32 << 20, (4 * n_params) // 12))
    reducer = BucketedGradientAllReduce(model.parameters(), bucket_bytes=bucket)
    b, e = shard_bounds(tcfg["B"], rank, world)
    gen = torch.Generator(device="cpu").manual_seed(5)
    x = torch.randn(tcfg["B"], tcfg["D"], generator=gen)[b:e].to(dev)
    weight = (e - b) * world / tcfg["B"]                    # unequal shards of a mean loss
    reducer.weight = float(weight)
    exposed = []

    def step():
        reducer.zero_grad()
        ll, _ = model.compute_ll(x)
        (-ll.mean()).backward()
        reducer.finish()
        opt.step()

    for _ in range(max(2, warmup)):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    s, ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        step()
        exposed.append(reducer._events)
    ev.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = s.elapsed_time(ev) / steps
    pairs = [p for p in exposed if p is not None]
    exp_ms = float(np.mean([a.elapsed_time(z) for a, z in pairs])) if pairs else 0.0
    if world > 1:
        t = torch.tensor([ms, exp_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, exp_ms = float(t[0]), float(t[1])
    out = {"flow": tcfg["label"], "global_batch": tcfg["B"], "per_gpu_batch": e - b, "ms_per_step": ms,
           "samples_per_s": tcfg["B"] / (ms * 1e-3), "allreduce_exposed_ms": exp_ms,
           "allreduce_bytes": reducer.total_bytes, "buckets": len(reducer.buckets), "bucket_bytes": bucket,
           "collective": f"NCCL all-reduce x{len(reducer.buckets)} buckets launched from gradient hooks" if world > 1 else "none (1 GPU)"}
    reducer.close()
    del model, opt, reducer
    torch.cuda.empty_cache()
    return out


def run_ours(args, cfg, name):
    import torch
    import torch.distributed as dist
    from oracle import c_binding, umnn_oracle as orc
    from umnn_b200 import IntegrandNN, IntegrandNetwork, _native, cc_integrate, cc_integrate_host
    from umnn_b200.distributed import shard_bounds

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
    Bg, D, E, Q = cfg["B"], cfg["D"], cfg["E"], cfg["Q"]
    lo, hi = shard_bounds(Bg, rank, world)              # STRONG scaling: this rank's shard of the global batch
    B = hi - lo
    Hh = E * D if cfg["layout"] == "strided" else E
    probe = _native.make_desc(_native.LAYOUT_STRIDED_D if cfg["layout"] == "strided" else _native.LAYOUT_CONTIG, max(B, 1), D, E,
                              spec.widths, _native.ACT_LEAKY_RELU if cfg["layout"] == "strided" else _native.ACT_RELU,
                              _native.OUT_ELU_PLUS_1, Q, _native.PREC_FP16X3)
    tc_ok = _native.lib().umnn_packed_params_bytes(probe) > 0
    use_tc = args.precision in ("bf16x3", "fp16x3") or (args.precision == "auto" and tc_ok)
    split = args.precision if args.precision != "auto" else "fp16x3"
    pad = lambda w: (w + 2 + 15) // 16 * 16
    issued_per_row = 3 * 2 * sum(pad(a) * pad(b) for a, b in zip(spec.widths[1:-2], spec.widths[2:-1])) if use_tc else 0
    net = IntegrandNetwork(D, 1 + E, cfg["hidden"], 1) if cfg["layout"] == "strided" else IntegrandNN(1 + E, cfg["hidden"])
    off = 0
    with torch.no_grad():
        for p in net.parameters():
            p.copy_(torch.from_numpy(flat[off:off + p.numel()].copy()).view_as(p))
            off += p.numel()
    net.to(dev).eval()     # inference-style step: packed parameters are cached

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

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step_resident()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_total = timed(step_resident, args.steps)
    clocks = sampler.stop() if sampler else None

    for _ in range(min(warm, 3)):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)          # the SAME K steps, through the host-buffer API

    # the same resident step through a prepared launch (umnn_b200.prepare_integral: descriptor, packed parameters and
    # argument checks resolved once) -- what a serving loop of one shape calls; matters where the host paces (config 1)
    from umnn_b200 import prepare_integral
    prep = prepare_integral(net, B, Q, want_fx=True)
    for _ in range(warm):
        prep(x, h)
    ms_prep = timed(lambda: prep(x, h), args.steps)

    # parity of what was just timed: >= 256 samples of this rank's shard against the C oracle -- contiguous runs at the
    # start (first CTA's slot range), the end (last CTA, ragged tail) and the middle of the shard, plus a random subset
    o, f, _ = step_resident()
    torch.cuda.synchronize()
    n_chk = min(B, 256 if D > 16 else 1024)
    run = max(1, n_chk // 4)
    picks = np.unique(np.concatenate([np.arange(0, run), np.arange(B - run, B), np.arange(max(0, B // 2 - run // 2), min(B, B // 2 + run - run // 2)),
                                      np.random.RandomState(rank).choice(B, min(B, run), replace=False)]))
    tp = torch.from_numpy(picks)
    ref, rfx, _ = c_binding.cc_forward(spec, flat, np.zeros((len(picks), D), np.float32), x_host[tp].numpy(),
                                       h_host[tp].numpy(), Q, cfg["layout"])
    got = o[tp.to(dev)].cpu().numpy()
    rel = float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-6)))
    jac_abs = float(np.max(np.abs(np.log(f[tp.to(dev)].cpu().numpy() + 1e-10) - np.log(rfx + 1e-10))))
    if world > 1:
        t = torch.tensor([rel, jac_abs], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        rel, jac_abs = float(t[0]), float(t[1])

    # training legs run on every rank (they contain the collective)
    train = {}
    if not args.no_train:
        for tname in ("cfg3", "cfg5"):
            try:
                train[tname] = train_step_leg(TRAIN_FLOWS[tname], dev, world, rank, steps=max(3, min(args.steps, 10)), warmup=2)
            except Exception as ex:      # a training leg must never take the headline line down with it
                train[tname] = {"error": f"{type(ex).__name__}: {ex}"[:300]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    evals_per_step = Bg * D * Q
    rows_per_step = Bg * D * (Q + 2)
    ms_per_step = ms_total / args.steps
    value = evals_per_step / (ms_per_step * 1e-3)
    e2e_value = evals_per_step / (ms_e2e / args.steps * 1e-3)
    peaks = load_peaks()
    fpe = flop_per_eval(spec.widths)
    # per-GPU kernel: rows of one launch x FLOP per row / average launch duration (one fused launch per step and GPU)
    rows_per_launch = B * D * (Q + 2)
    achieved_tflops = rows_per_launch * fpe / (ms_per_step * 1e-3) / 1e12
    peak = peaks["bf16_tflops_sustained"]
    slot_bytes = 4 * (E + 1) + 8            # x, h read (x0 = NULL); integral, f(x) written
    traffic, traffic_src = None, None
    tpath = os.path.join(REPO, "profiles", "traffic.json")
    if os.path.exists(tpath):
        rec = json.load(open(tpath)).get(f"{name}:{'tc' if use_tc else 'fp32'}")
        if rec:
            traffic = rec["dram_bytes_per_slot"] * B * D
            traffic_src = ("replayed from an ncu capture, NOT measured in this run: dram__bytes_read.sum + dram__bytes_write.sum of one "
                           f"launch of this kernel ({rec.get('source', 'profiles/')}), recorded per slot and scaled to this batch")
    out = {
        "metric": "integrand-evals/sec (B*D*Q)", "value": value, "unit": "integrand-evals/s", "n_gpus": world,
        "steps": args.steps, "warmup": warm, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None,
        "dtype": f"{split} (16-bit hi+lo operand split, fp32 accumulate)" if use_tc else "f32",
        "data": "synthetic",
        "config": {"workload": f"{name}: {cfg['label']}", "global_batch": Bg, "per_gpu_batch": B, "D": D,
                   "E": E, "Q": Q, "hidden": cfg["hidden"],
                   "parallelism": f"global batch sharded over {world} GPU(s) ({B} samples on rank 0), no data-path collective",
                   "rows_per_step": rows_per_step, "precision": args.precision,
                   "kernel": "cc_forward_tc (tcgen05 cta_group::2)" if use_tc else "cc_forward_fp32 (FFMA)",
                   "l2": f"inputs {(x.numel() + h.numel()) * 4 / 1e6:.0f} MB per step and GPU "
                         + ("> 126 MB L2 (no flush needed)" if (x.numel() + h.numel()) * 4 > 126e6 else "< L2: resident")},
        "e2e": {"value": e2e_value, "unit": "integrand-evals/s", "ms_per_step": ms_e2e / args.steps, "steps": args.steps,
                "h2d_bytes_per_step": int((x_host.numel() + h_host.numel()) * 4) * world,
                "d2h_bytes_per_step": int((out_host.numel() + fx_host.numel()) * 4) * world,
                "api": "umnn_b200.cc_integrate_host on pinned host tensors (H2D + fused kernel + D2H every step, "
                       "pipelined over batch chunks)"},
        # fp16x3: the fused kernel + its guarded FP32 re-run (a no-op launch unless an activation left the fp16 range)
        "gpu_launches": args.steps * world * (2 if (use_tc and split == "fp16x3") else 1),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "achieved": achieved_tflops, "peak": peak, "unit": "TFLOP/s",
                     "frac": achieved_tflops / peak, "traffic": traffic, "traffic_source": traffic_src,
                     "algorithmic_bytes": B * D * slot_bytes,
                     "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peaks['source']})",
                     "flop_per_row": fpe, "rows_per_launch": rows_per_launch,
                     "hbm_sanity_gbs": (B * D * slot_bytes) / (ms_per_step * 1e-3) / 1e9,
                     "issued_tensor_tflops": rows_per_launch * issued_per_row / (ms_per_step * 1e-3) / 1e12,
                     "issued_tensor_frac": rows_per_launch * issued_per_row / (ms_per_step * 1e-3) / 1e12 / peak,
                     "note": "per GPU.  achieved = algorithmic fp32 FLOP (2*sum in*out per row) / time; the hi+lo operand split issues 3 "
                             "padded 16-bit MMAs per algorithmic MAC on the hidden layers, so frac is capped near 1/3.3; "
                             "issued_tensor_* counts the 16-bit FLOP actually sent to the tensor pipe"},
        "parity": {"integral_max_rel_err_vs_oracle": rel, "log_jac_max_abs_err_vs_oracle": jac_abs, "samples_per_rank": int(len(picks)),
                   "where": "contiguous runs at the start, middle and end of each rank's shard + a random subset, C oracle"},
    }
    aux = {"prepared_call": {"ms_per_step": ms_prep / args.steps, "value": evals_per_step / (ms_prep / args.steps * 1e-3),
                             "api": "umnn_b200.prepare_integral(net, B, Q, want_fx=True)(x, h), inputs resident"}}
    if train:
        aux["train_step"] = train
    if world == 1 and not args.no_cpu:
        refcpu = reference_cpu_arm(cfg, flat, budget_s=12.0)
        port = cpu_port_arm(cfg, spec, flat)
        if refcpu is not None:
            v, info = refcpu
            out["cpu_baseline"] = {"value": v, "unit": "integrand-evals/s", "cores": info["cores"], "kind": "reference",
                                   "sample": info["sample"]}
            aux["cpu_port"] = port
        else:
            out["cpu_baseline"] = port
        aux["reference_torch_cuda"] = reference_cuda_probe(cfg, flat, dev)
        if "value" in aux["reference_torch_cuda"]:
            aux["speedup_vs_reference_torch_cuda"] = value / aux["reference_torch_cuda"]["value"]
        aux["backward"] = backward_probe(net, cfg, dev)
    if aux:
        out["aux"] = aux
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
    ap.add_argument("--batch", type=int, default=0, help="override the global batch (debugging only)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline / reference-CUDA / backward probes")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step legs")
    ap.add_argument("--precision", default=os.environ.get("UMNN_B200_PRECISION", "auto"), choices=["auto", "fp32", "bf16x3", "fp16x3"],
                    help="kernel family: auto = tensor cores (fp16x3 split, guarded FP32 re-run) when the shape fits, else FP32 FFMA")
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
