#!/usr/bin/env python
"""bench.py -- IAF-transform throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c2a|c2b]

A "step" is one fused IAF step (masked-AR conv stack -> mu, s -> z' = (z - .1 mu)/exp(.1 s),
per-element arw_logsd, per-sample logdet) over one batch of 256 synthetic samples of
n_z=32, 16x16.  Metric: latent elements/s = B*n_z*H*W / t_step, whole job.

* value      : inputs resident in HBM, K steps timed with CUDA events (one CUDA graph of K
               launches, or K direct launches), max over ranks.  The K steps rotate through
               NSETS input/output sets whose footprint exceeds L2, so no step finds its
               inputs in L2.
* e2e        : same metric through the public host-buffer entry (IAFOperator.submit_host ->
               iaf_step_submit_host): pinned host inputs H2D, step, results D2H, every step,
               pipelined over three device staging slots; timed until wait_host() returns.
* roofline   : the step kernel against the measured HBM (or bf16 tensor) peak.
* cpu_baseline: the oracle's torch-CPU port of the reference path on this box's cores,
               on a bounded sample of the same workload (rank 0, N=1).
* --impl reference: times that CPU port instead (the reference's Theano/TF code cannot
               run in this image), same JSON shape.

N>1: launched by torchrun, one rank per GPU; the batch dimension is sharded (every rank
owns 256 samples: weak scaling), the only collective is one NCCL all-reduce of the scalar
sum of log-dets (the ELBO term) at the end of the timed region (tf_train.py:142).
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

WORKLOADS = {
    # name: (variant, n_z, hidden, H, W, B, roofline bound)
    "c2a": ("tf", 32, [64], 16, 16, 256, "hbm"),
    "c2b": ("tf", 32, [160, 160], 16, 16, 256, "tensor"),
    # per-step shapes of the other BASELINE configs (parity-test cases; benched for the record, not the headline)
    "c1": ("theano", 32, [64], 16, 16, 16, "hbm"),              # README example, batch 16, level 0
    "c1_l1": ("theano", 32, [64], 8, 8, 16, "hbm"),             # ... level 1
    "c1_l2": ("theano", 32, [64], 4, 4, 16, "hbm"),             # ... level 2
    "c3": ("tf", 32, [160, 160], 16, 16, 32, "tensor"),         # tf_train.py default per-GPU batch
    "c4_l1": ("theano", 32, [160, 160], 8, 8, 16, "tensor"),    # Table-3 config, second level
}
METRIC = "IAF latents/sec (z',logdet) @ n_z=32,16x16,bs256"
UNIT = "latent elements/s"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


class ClockSampler(object):
    """Polls SM clock / throttle reasons through NVML while the timed regions run."""

    def __init__(self, index):
        self.samples = []
        self.phase = "idle"
        self.stop = False
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = repr(e)
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        while not self.stop:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((self.phase, mhz, reasons))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.ok:
            self.t.start()

    def finish(self):
        self.stop = True
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "NVML unavailable"}
        self.t.join(timeout=1.0)
        names = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
                 0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
                 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}
        timed = [s for s in self.samples if s[0] == "timed"]
        window = "timed region"
        if len(timed) < 3:
            timed = [s for s in self.samples if s[0] in ("timed", "e2e", "warmup")]
            window = "warmup+timed+e2e (timed region shorter than 3 samples)"
        if not timed:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "no samples"}
        bits = 0
        for s in timed:
            bits |= s[2]
        reasons = [n for b, n in names.items() if bits & b and n != "gpu_idle"]
        return {"sm_mhz": statistics.median(s[1] for s in timed), "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(timed), "window": window}


def workload_string(name):
    variant, n_z, hidden, H, W, B, _ = WORKLOADS[name]
    return "%s: single IAF step, n_z=%d hidden=%s %dx%d batch %d per GPU, %s-variant numerics" % (
        name, n_z, hidden, H, W, B, variant)


def make_layers(name, seed=0):
    """Seeded synthetic (direction, gain, bias) per conv in the variant's own layout: TF V[3,3,Cin,Cout], g, b
    (layers.py:53-55); Theano w[Cout,Cin+1,3,3], s, b (ar.py:288-296; its gain is exp(3 s))."""
    variant, n_z, hidden = WORKLOADS[name][:3]
    g = torch.Generator().manual_seed(seed + 1)
    sizes = [n_z] + hidden
    layers = []
    for i in range(len(hidden) + 2):
        cin = sizes[min(i, len(hidden))]
        cout = hidden[i] if i < len(hidden) else n_z
        shape = (3, 3, cin, cout) if variant == "tf" else (cout, cin + 1, 3, 3)
        V = 0.05 * torch.randn(shape, generator=g)
        gg = torch.rand((cout,), generator=g) - 0.5
        if variant == "theano":
            gg = gg / 3.0
        b = 0.1 * torch.randn((cout,), generator=g)
        layers.append((V, gg, b))
    return layers


def make_workload(name, device, nsets, seed=0):
    from iaf_b200 import IAFOperator
    variant, n_z, hidden, H, W, B, bound = WORKLOADS[name]
    layers = make_layers(name, seed)
    op = IAFOperator(variant, n_z, hidden, [n_z, n_z], nl="elu", path="auto")
    op.set_weights([tuple(t.to(device) for t in l) for l in layers])
    g = torch.Generator().manual_seed(seed)
    sets = []
    logdets = torch.zeros((nsets, B), device=device)  # one row per set: the ELBO scalar is one .sum() over it
    for i in range(nsets):
        z = torch.randn((B, n_z, H, W), generator=g)
        ctx = 0.1 * torch.randn((B, hidden[0], H, W), generator=g)
        sets.append(dict(z=z.to(device), ctx=ctx.to(device), z_out=torch.empty((B, n_z, H, W), device=device),
                         logsd=torch.empty((B, n_z, H, W), device=device), logdet=logdets[i]))
    op.logdets = logdets
    return op, layers, sets


def cpu_port_runner(name, layers_cpu, sample_B, threads):
    """Returns (fn, elems_per_call): one reference-path IAF step on the host cores."""
    from oracle import iaf_oracle_torch as OT
    variant, n_z, hidden, H, W, B, _ = WORKLOADS[name]
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(0)
    z = torch.randn((sample_B, n_z, H, W), generator=g)
    ctx = 0.1 * torch.randn((sample_B, hidden[0], H, W), generator=g)
    keys = ("V", "g", "b") if variant == "tf" else ("w", "s", "b")
    hid = [dict(zip(keys, l)) for l in layers_cpu[:len(hidden)]]
    heads = [dict(zip(keys, l)) for l in layers_cpu[len(hidden):]]

    def fn():
        with torch.no_grad():
            return OT.iaf_step(variant, z, ctx, hid, heads, "elu")
    return fn, sample_B * n_z * H * W


def best_thread_count(fn, max_threads):
    """The reference's framework would pick its own thread count; more threads than the convs can
    use makes torch slower, so probe a few settings and keep the fastest."""
    best, best_t = max_threads, None
    cands = sorted({c for c in (8, 16, 32, 64, max_threads) if c <= max_threads})
    for c in cands:
        torch.set_num_threads(c)
        fn()
        t0 = time.perf_counter()
        fn()
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = c, dt
    torch.set_num_threads(best)
    return best


def time_cpu(fn, warmup, steps):
    for _ in range(warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    return (time.perf_counter() - t0) / steps


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path (torch-CPU port of the oracle), rank 0 only."""
    if rank != 0:
        return
    name = args.workload
    variant, n_z, hidden, H, W, B, bound = WORKLOADS[name]
    threads = os.cpu_count() or 1
    layers = make_layers(name)
    # bounded sample: the full 256-sample batch per step, at most 50 steps
    sample_B = B
    fn, elems = cpu_port_runner(name, layers, sample_B, threads)
    threads = best_thread_count(fn, threads)
    steps = max(1, min(args.steps, 50))
    warm = max(1, min(args.warmup, 3))
    t = time_cpu(fn, warm, steps)
    value = elems / t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": t * 1e3 * (B / sample_B), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_string(name), "sample": "%d of %d samples per step, %d steps" % (sample_B, B, steps)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d of the %d samples per step, %d steps, torch-CPU fp32 port of the "
                                   "reference path (Theano/TF originals cannot run here)" % (sample_B, B, steps)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2a", choices=sorted(WORKLOADS))  # c2a = the headline
    ap.add_argument("--no-graph", action="store_true", help="K direct launches instead of one CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import __graft_entry__
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    if rank == 0:
        __graft_entry__.build()
    if dist is not None:
        dist.barrier()

    name = args.workload
    variant, n_z, hidden, H, W, B, bound = WORKLOADS[name]
    K, Wm = args.steps, args.warmup
    alg_bytes_unit = 4 * B * H * W * (n_z + hidden[0] + n_z + n_z) + 4 * B
    nsets = max(2, -(-3 * 126 * 2 ** 20 // alg_bytes_unit))  # footprint >= 3x the 126 MB L2
    op, layers_cpu, sets = make_workload(name, device, nsets)
    lib = op._lib
    import ctypes as C
    plan = op._plan(H, W, device)
    stream = torch.cuda.current_stream(device)

    def launch(i, st):
        s = sets[i % nsets]
        rc = lib.iaf_step_fwd(plan, C.c_void_p(s["z"].data_ptr()), C.c_void_p(s["ctx"].data_ptr()),
                              C.c_void_p(s["z_out"].data_ptr()), C.c_void_p(s["logsd"].data_ptr()),
                              C.c_void_p(s["logdet"].data_ptr()), B, C.c_void_p(st.cuda_stream))
        if rc != 0:
            from iaf_b200 import _lib
            _lib.check(rc)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        sampler.phase = "warmup"
    lc0 = op.launch_count()
    for i in range(Wm):
        launch(i, stream)
    torch.cuda.synchronize()
    launches_per_step = (op.launch_count() - lc0) // Wm  # 1 (fused / SIMT kernel) or one per conv stage (layered)

    graph = None
    launch_mode = "direct"
    if not args.no_graph:
        try:
            gstream = torch.cuda.Stream(device)
            gstream.wait_stream(stream)
            with torch.cuda.stream(gstream):
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=gstream):
                    for i in range(K):
                        launch(i, torch.cuda.current_stream(device))
            stream.wait_stream(gstream)
            launch_mode = "cuda_graph(%d launches)" % K
        except Exception as e:  # capture unsupported -> direct launches (still the CUDA path)
            graph = None
            launch_mode = "direct (graph capture failed: %s)" % type(e).__name__
            torch.cuda.synchronize()
    if graph is not None:
        graph.replay()  # one untimed replay
        torch.cuda.synchronize()

    # ---- timed region: device-resident ----
    launches0 = op.launch_count()
    ev0, ev1, ev2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    total = op.logdets.sum()          # warm the reduction (and the collective) outside the timed region
    if dist is not None:
        dist.all_reduce(total)
        dist.barrier()
    torch.cuda.synchronize()
    if sampler:
        sampler.phase = "timed"
    ev0.record()
    if graph is not None:
        graph.replay()
    else:
        for i in range(K):
            launch(i, stream)
    ev1.record()
    # the ELBO scalar: sum of log-dets of the sets touched, one all-reduce (tf_train.py:142)
    total = op.logdets.sum()
    if dist is not None:
        dist.all_reduce(total)
    ev2.record()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    if sampler:
        sampler.phase = "between"
    n_launched = K * launches_per_step if graph is not None else op.launch_count() - launches0
    t_kernels_ms = ev0.elapsed_time(ev1)
    t_total_ms = ev0.elapsed_time(ev2)
    tt = torch.tensor([t_total_ms, t_kernels_ms], device=device, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_total_ms, t_kernels_ms = float(tt[0]), float(tt[1])
    elems_step = B * n_z * H * W
    value = world * elems_step * K / (t_total_ms * 1e-3)

    # ---- e2e: host buffers through the public API ----
    # every step: pinned host inputs -> H2D -> step -> D2H of z', arw_logsd, logdet into pinned host outputs.
    # Steps go through the pipelined public entry (iaf_step_submit_host): three staging slots, so the copy-in
    # of step i+1, the kernel of step i and the copy-out of step i-1 overlap; the region ends after wait_host().
    NH = 4
    hz = [torch.empty((B, n_z, H, W)).pin_memory().copy_(sets[i % nsets]["z"].cpu()) for i in range(NH)]
    hc = [torch.empty((B, hidden[0], H, W)).pin_memory().copy_(sets[i % nsets]["ctx"].cpu()) for i in range(NH)]
    ho = [torch.empty((B, n_z, H, W)).pin_memory() for _ in range(NH)]
    hl = [torch.empty((B, n_z, H, W)).pin_memory() for _ in range(NH)]
    hd = [torch.empty((B,)).pin_memory() for _ in range(NH)]
    Ke = K
    for i in range(4):
        op.submit_host(hz[i % NH], hc[i % NH], ho[i % NH], hl[i % NH], hd[i % NH])
    op.wait_host()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    if sampler:
        sampler.phase = "e2e"
    t0 = time.perf_counter()
    for i in range(Ke):
        op.submit_host(hz[i % NH], hc[i % NH], ho[i % NH], hl[i % NH], hd[i % NH])
    op.wait_host()
    t_e2e = time.perf_counter() - t0
    e2e_check = float(hd[(Ke - 1) % NH].sum())  # the step's result is read on the host
    te = torch.tensor([t_e2e], device=device, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    t_e2e = float(te[0])
    e2e_value = world * elems_step * Ke / t_e2e
    h2d = hz[0].numel() * 4 + hc[0].numel() * 4
    d2h = ho[0].numel() * 4 + hl[0].numel() * 4 + hd[0].numel() * 4
    clocks = sampler.finish() if sampler else None

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the step kernel ----
    hbm_gbs, bf16_tf, peak_src = measured_peaks()
    t_kernel = t_kernels_ms * 1e-3 / K
    alg_bytes = op.algorithmic_bytes(B, H, W, device)
    alg_flops = op.algorithmic_flops(B, H, W, device)
    traffic = None
    tj = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tj):
        with open(tj) as f:
            traffic = json.load(f).get(name + ":" + op.path_used(H, W, device))
    if bound == "hbm":
        achieved = alg_bytes / t_kernel / 1e9
        roof = {"bound": "hbm", "achieved": achieved, "peak": hbm_gbs, "unit": "GB/s", "frac": achieved / hbm_gbs}
    else:
        achieved = alg_flops / t_kernel / 1e12
        roof = {"bound": "tensor", "achieved": achieved, "peak": bf16_tf, "unit": "TFLOP/s", "frac": achieved / bf16_tf}
    roof.update({"traffic": traffic, "kernel": "iaf_step (%s path)" % op.path_used(H, W, device),
                 "kernel_us": t_kernel * 1e6, "algorithmic_bytes": alg_bytes, "algorithmic_flops": alg_flops,
                 "peak_source": peak_src})

    # ---- CPU baseline (bounded sample) ----
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        threads = os.cpu_count() or 1
        sample_B = B
        fn, elems = cpu_port_runner(name, layers_cpu, sample_B, threads)
        threads = best_thread_count(fn, threads)
        t1 = time_cpu(fn, 2, 1)
        reps = int(max(3, min(200, 15.0 / max(t1, 1e-4))))
        t = time_cpu(fn, 0, reps)
        cpu = {"value": elems / t, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "%d of the %d samples per step x %d steps (%.1f s), torch-CPU fp32 port of the "
                         "reference path" % (sample_B, B, reps, t * reps)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": t_total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 (tc path: bf16x3 split operands, f32 accumulate)" if op.path_used(H, W, device) == "tc" else "f32",
        "data": "synthetic",
        "config": {"workload": workload_string(name), "global_batch": B * world, "parallelism": "dp%d" % world,
                   "path": op.path_used(H, W, device), "launch": launch_mode, "kernels_per_step": launches_per_step,
                   "l2": "rotating %d input/output sets (%.0f MB > 126 MB L2)" % (nsets, nsets * alg_bytes_unit / 2 ** 20),
                   "collective": "one all-reduce of the scalar sum(logdet) per timed region" if world > 1 else "none",
                   "samples_per_s": value / (n_z * H * W)},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": t_e2e / Ke * 1e3, "steps": Ke, "logdet_sum_last_step": e2e_check,
                "entry": "IAFOperator.submit_host/wait_host -> iaf_step_submit_host (3-slot H2D/compute/D2H pipeline)"},
        "gpu_launches": int(n_launched),
        "clocks": clocks,
        "roofline": roof,
        "cpu_baseline": cpu,
        "elbo_scalar": float(total),
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
