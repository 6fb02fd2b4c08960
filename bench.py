#!/usr/bin/env python
"""Benchmark of the KAGNN hot path on B200 (contract: see the task statement / DESIGN.md section "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload = BASELINE.json configs[1]: ogbn-arxiv-shaped KAGIN -- ``GKAN_Nodes('gin', 3, 128, 64, 40, skip=True,
grid_size=5, spline_order=3, hidden_layers=2)`` in eval mode, N = 169 343 nodes, E = 1 166 243 directed edges
(uniform random pairs, kept directed), x ~ N(0, 0.3^2), fp32.  A "step" is one model forward over the whole graph.
With N > 1 GPUs every rank owns an arxiv-sized node range of an N-times larger random graph (weak scaling) and
exchanges halo rows once per layer (kagnn_b200/dist.py).

Prints ONE JSON line.  ``value`` = nodes/s with inputs resident in HBM (CSR cached, as in the reference's training
loop where the same edge_index is reused every epoch); ``e2e`` = the same through the public module API with pinned
HOST buffers: H2D of x and edge_index, CSR build, forward, D2H of the logits, all inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

N_NODES, N_EDGES, N_FEAT, HIDDEN, N_CLASSES, MP_LAYERS, KAN_DEPTH, GRID, ORDER = 169_343, 1_166_243, 128, 64, 40, 3, 2, 5, 3
METRIC = "KAGNN-layer forward nodes/sec"
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture
# (profiles/r2_ncu_full_summary.csv: 186.43+28.46 MB, 50.56+8.02 MB, 217.40+15.70 MB)
NCU_TRAFFIC_BYTES = {"agg1[128]->64->64": 214_891_776, "agg1[64]->64->64": 58_577_664, "agg0[320]->40": 233_096_960}
WORKLOAD = "ogbn-arxiv-shaped KAGIN (GKAN_Nodes gin, 3 layers, hidden 64, grid 5, order 3, KAN depth 2), fp32, eval"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return json.load(fh), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def synth_graph(n, e, f, seed, n_src=None):
    g = torch.Generator().manual_seed(seed)
    src = torch.randint(0, n_src or n, (e,), generator=g)
    dst = torch.randint(0, n, (e,), generator=g)
    x = torch.randn(n, f, generator=g) * 0.3
    return x, torch.stack([src, dst])


def model_state(seed=12345):
    """Random-init weights of the architecture (no checkpoints offline) + BN statistics that keep hidden
    activations O(1), built once on CPU so both arms use the same numbers."""
    import kagnn_b200 as kb
    torch.manual_seed(seed)
    m = kb.GKAN_Nodes("gin", MP_LAYERS, N_FEAT, HIDDEN, N_CLASSES, skip=True, grid_size=GRID, spline_order=ORDER,
                      hidden_layers=KAN_DEPTH, dropout=0.0).eval()
    with torch.no_grad():
        for name, b in m.named_buffers():
            if name.endswith("running_var"):
                b.fill_(0.02)
    return m


PRE_WARM_S = 0.3
PRE_WARM_STEPS = 200                # the same phase as a step count, for sharded runs


class ClockSampler:
    """SM clock / throttle reasons DURING the timed region: an NVML polling thread (2 ms period; the timed region of this
    bench is tens of milliseconds, too short for `nvidia-smi -lms`), falling back to the nvidia-smi query of
    B200_PROFILING.md when NVML cannot be loaded."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.samples, self.stop_flag, self.proc, self.nvml, self.index = [], False, None, None, index
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.th = threading.Thread(target=self._poll, daemon=True)
            self.th.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nvml
        names = (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap))
        while not self.stop_flag:
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            except Exception:
                time.sleep(0.002)
                continue
            mask = None                                   # the reasons query is not available on every driver / permission level:
            for fn in ("nvmlDeviceGetCurrentClocksEventReasons", "nvmlDeviceGetCurrentClocksThrottleReasons"):
                try:                                      # a clock sample must not be lost with it
                    mask = getattr(nv, fn)(self.h)
                    break
                except Exception:
                    continue
            reasons = ["reasons unavailable"] if mask is None else [n for n, b in names if mask & b]
            self.samples.append((time.perf_counter(), mhz, self.max, reasons))
            time.sleep(0.002)

    def _read(self):
        for ln in self.proc.stdout:
            parts = [q.strip() for q in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                mhz, mx = float(parts[0]), float(parts[1])
            except ValueError:
                continue
            rs = [n for n, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7])
                  if v.lower().startswith("active")]
            self.samples.append((time.perf_counter(), mhz, mx, rs))

    def stop(self, t0=None, t1=None):
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()
        if self.nvml is None and self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"], "samples": 0}
        time.sleep(0.01)
        sel = [q for q in self.samples if t0 is None or (t0 <= q[0] <= t1)]
        if not sel:                                       # region shorter than one period: nearest sample
            sel = sorted(self.samples, key=lambda q: abs(q[0] - (t0 or 0)))[:1]
        if not sel:                                       # the poller got nothing at all: one query now, the GPU is still warm
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=10).stdout.strip().splitlines()[0]
                parts = [q.strip() for q in out.split(",")]
                rs = [n for n, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7])
                      if v.lower().startswith("active")]
                return {"sm_mhz": float(parts[0]), "sm_max_mhz": float(parts[1]), "reasons": rs, "samples": 1,
                        "source": "nvidia-smi, one query right after the timed region (the poller returned nothing)"}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"], "samples": 0}
        reasons = sorted({r for q in sel for r in q[3]})
        return {"sm_mhz": statistics.median([q[1] for q in sel]) if sel else None, "sm_max_mhz": sel[0][2] if sel else None,
                "reasons": reasons, "samples": len(sel), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def layer_algorithmic_bytes(n, e_agg, f_in, f_out, params, self_rows=True, per_edge_scalar=False):
    """fp32, int32 CSR, no-reuse gather model."""
    b = 4 * f_in * e_agg + 4 * e_agg + 4 * (n + 1) + 4 * f_out * n + 4 * params
    if self_rows:
        b += 4 * f_in * n
    if per_edge_scalar:
        b += 4 * e_agg
    return b


def kan_params(sizes, S):
    return sum(i * o * (S + 2) for i, o in zip(sizes[:-1], sizes[1:]))     # spline (S) + base + scaler per (in,out) pair


def cpu_baseline_sample(sd, frac=1, threads=None, steps=1, warmup=1):
    """The oracle (torch-CPU restatement of the reference forward) on the workload's graph (frac = 1: the full N = 169 343
    nodes / 1 166 243 edges, one forward takes a few seconds on the box's host cores)."""
    from oracle import kagnn_oracle as K
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    n, e = N_NODES // frac, N_EDGES // frac
    x, ei = synth_graph(n, e, N_FEAT, 777)
    ts = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            K.node_model_forward(sd, "gin", x, ei, True)
            if i >= warmup:
                ts.append(time.perf_counter() - t0)
    return n, e, ts, threads


def run_reference(args, rank):
    """--impl reference: the reference's own PyTorch-CPU forward (restated: torch_geometric cannot be installed),
    all host threads, each step = one forward of the same model on the FULL workload graph (same config as the GPU arm)."""
    if rank != 0:
        return
    m = model_state()
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    frac = 1
    # a full-size step takes ~6 s on the box's host cores: K + W up to 40 steps stays within a few minutes; beyond that the run
    # is bounded (the line reports the steps that were actually timed)
    if args.steps + args.warmup > 40:
        args.warmup = min(args.warmup, 5)
        args.steps = min(args.steps, 35)
    n, e, ts, threads = cpu_baseline_sample(sd, frac, steps=args.steps, warmup=args.warmup)
    total = sum(ts)
    value = n * len(ts) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "nodes/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(ts), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "nodes": N_NODES, "edges": N_EDGES, "features": N_FEAT,
                   "nodes_per_gpu": N_NODES, "edges_per_gpu": N_EDGES,
                   "note": "reference arm = oracle port of the reference forward on host cores (torch_geometric not installable)"},
        "cpu_baseline": {"value": value, "unit": "nodes/s", "cores": threads, "kind": "port",
                         "sample": f"full model forward on the full {n}-node / {e}-edge workload graph, {len(ts)} steps"},
        "e2e": {"value": value, "unit": "nodes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary configurations (key 'extra' of the line)")
    ap.add_argument("--config", default="arxiv", choices=["arxiv", "cora", "zinc", "mutag", "rmat"],
                    help="arxiv = the headline workload (BASELINE configs[1]); the others print the secondary measurement of that "
                         "configuration alone (scripts/bench_extras.py) as one JSON line")
    ap.add_argument("--resident-x-halo", action="store_true",
                    help="N > 1: keep the halo rows of the static input features between steps (part of the partitioned input); "
                         "default off = the input halo crosses NVLink every step")
    ap.add_argument("--dist-mode", default="auto", choices=["auto", "peer", "pull", "pull_overlap", "push", "halo"],
                    help="N > 1: 'peer' = in-kernel NVLink gather from symmetric memory, 'halo' = NCCL all-to-all per layer")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)            # timing rules: at least 3 warm-up steps (the line reports what was run)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the kagnn_b200 path has no CPU fallback")
    if args.config != "arxiv":
        # The other BASELINE configurations.  Graph-level batches (zinc, mutag) shard by graph with no exchange: under torchrun
        # every rank runs its own batch and the line reports the aggregate (max over ranks of the step time); the node-level
        # ones (cora, rmat) run on rank 0 only.
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        import bench_extras as BX
        sharded = world > 1 and args.config in ("zinc", "mutag")
        if rank != 0 and not sharded:
            return
        torch.cuda.set_device(local_rank)
        dev = torch.device("cuda", local_rank)
        if sharded:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=dev)
        flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
        fn = getattr(BX, args.config)
        res = fn(dev, flush) if args.config == "rmat" else fn(dev, flush, with_cpu=(not args.no_cpu_baseline) and rank == 0)
        ms = res.get("ms", res.get("ms_layer", res.get("ms_fp32")))
        nodes = {"cora": 2708}.get(args.config, res.get("nodes"))
        n_ranks = 1
        if sharded:
            t = torch.tensor([ms, float(res.get("ms_bf16", 0.0))], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, n_ranks = float(t[0]), world
            res["ms_bf16_max_over_ranks"] = float(t[1])
            dist.barrier()
            dist.destroy_process_group()
        if rank == 0:
            # the same line shape as the headline run (metric = forward nodes/s of that configuration's model or layer)
            line = {"metric": METRIC, "value": n_ranks * nodes / ms * 1e3 if (ms and nodes) else None, "unit": "nodes/s", "n_gpus": n_ranks,
                    "steps": 20 if args.config == "cora" else 10, "warmup": 3, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                    "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                    "config": {"workload": res.get("workload", args.config), "l2": "flushed between iterations (512 MB memset)",
                               "parallelism": (f"graph batches sharded by graph x{n_ranks}, no exchange" if sharded else "single GPU")},
                    "e2e": ({"value": nodes / res["ms_e2e"] * 1e3, "unit": "nodes/s", "ms_per_step": res["ms_e2e"]} if "ms_e2e" in res else None),
                    "cpu_baseline": ({"value": nodes / res["cpu_port_ms"] * 1e3, "unit": "nodes/s", "cores": res.get("cpu_threads"), "kind": "port",
                                      "sample": "oracle forward of the same model on the same inputs"} if "cpu_port_ms" in res else None),
                    "roofline": res.get("roofline"), "detail": {k: v for k, v in res.items() if k not in ("workload", "roofline")}}
            print(json.dumps(line), flush=True)
        return
    import kagnn_b200 as kb
    from kagnn_b200 import ops

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist_on = world > 1
    if dist_on:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    pk, pk_src = peaks()

    model = model_state()
    sd_cpu = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = model.to(dev)
    n_local = N_NODES
    x_host, ei_host = synth_graph(n_local, N_EDGES, N_FEAT, 12345 + rank, n_src=n_local * world)
    x_host, ei_host = x_host.pin_memory(), ei_host.pin_memory()

    if dist_on:
        from kagnn_b200 import dist as kdist
        runner = kdist.ShardedNodeModel(model, rank, world, n_local, mode=args.dist_mode, resident_x_halo=args.resident_x_halo)
        ei_glob = ei_host.to(dev)
        ei_glob[1] += rank * n_local                                     # targets: this rank's node range, global ids
        plan = runner.prepare(ei_glob)
        in_symm = runner.mode in ("peer", "pull", "pull_overlap", "push")
        if in_symm:
            x_dev = runner.input_buffer(N_FEAT, dev)                     # x resident where the peers can read it (symmetric memory)
            x_dev.copy_(x_host)
        else:
            x_dev = x_host.to(dev)
        step = lambda: runner.forward(x_dev, plan)                       # noqa: E731
    else:
        x_dev, ei_dev = x_host.to(dev), ei_host.to(dev)
        step = lambda: model(x_dev, ei_dev)                              # noqa: E731
    flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # 512 MB > 126 MB L2

    def barrier():
        if dist_on:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        # a GPU that has been idle (a fresh box) takes a few hundred milliseconds of load to reach its boost clock: run untimed steps
        # for PRE_WARM_S first, then the W warm-up steps the command line asks for
        t_pre, n_pre = time.perf_counter(), 0
        while (n_pre < PRE_WARM_STEPS) if dist_on else (time.perf_counter() - t_pre < PRE_WARM_S):
            flush.zero_()                                 # (sharded: a fixed count -- every rank must take part in every step's barriers)
            step()
            torch.cuda.synchronize()
            n_pre += 1
        for _ in range(args.warmup):
            flush.zero_()
            step()
        barrier()
        sampler = ClockSampler(local_rank)
        evs = []
        l0 = ops.launch_count
        t_begin = time.perf_counter()
        for _ in range(args.steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step()
            e1.record()
            evs.append((e0, e1))
        barrier()
        t_end = time.perf_counter()
        launches = ops.launch_count - l0
        clocks = sampler.stop(t_begin, t_end)
        ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if dist_on:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())

        # ---- per-kernel timing of the same step (roofline of the dominant kernel) -----------------------
        per = {}
        for it in range(3 + 5):
            flush.zero_()
            ops.fused_timing = []
            step()
            torch.cuda.synchronize()
            if it >= 3:
                for idx, (label, a, b) in enumerate(ops.fused_timing):
                    per.setdefault((idx, label), []).append(a.elapsed_time(b))
            ops.fused_timing = None
        kern = sorted(((statistics.mean(v), k) for k, v in per.items()), reverse=True)
        kernels = [{"launch": k[0], "label": k[1], "ms": round(m_, 4)} for m_, k in kern]

        # ---- e2e: pinned host buffers in, logits out, through the module API ----------------------------
        # Per step, inside the timed region: H2D of edge_index (int64 COO) and of x from pinned host memory, CSR build
        # (kagnn_b200.graph.get_graph; it overlaps the x copy, which runs on a second stream), halo plan + exchanges when
        # sharded, the model forward, and D2H of the logits into a pinned host buffer.
        from kagnn_b200.graph import get_graph
        copy_stream = torch.cuda.Stream(device=dev)
        y_host = torch.empty(n_local, N_CLASSES, dtype=torch.float32).pin_memory()

        def e2e_step():
            cur = torch.cuda.current_stream()
            copy_stream.wait_stream(cur)
            with torch.cuda.stream(copy_stream):
                if dist_on and in_symm:
                    xd = runner.input_buffer(N_FEAT, dev)     # H2D straight into the symmetric buffer the peers read
                    xd.copy_(x_host, non_blocking=True)
                else:
                    xd = x_host.to(dev, non_blocking=True)
            ed = ei_host.to(dev, non_blocking=True)
            if dist_on:
                ed[1] += rank * n_local
                pl = runner.prepare(ed)                  # halo plan (index all-to-all) + shard CSR
                cur.wait_stream(copy_stream)
                if not in_symm:
                    xd.record_stream(cur)
                y = runner.forward(xd, pl)
            else:
                get_graph(ed, n_local)                   # COO -> CSR on the compute stream while x is still in flight
                cur.wait_stream(copy_stream)
                xd.record_stream(cur)
                y = model(xd, ed)
            y_host.copy_(y, non_blocking=True)

        for _ in range(3):
            flush.zero_()
            e2e_step()
        barrier()
        ts = []
        for _ in range(max(5, min(args.steps, 20))):
            flush.zero_()
            barrier()
            t0 = time.perf_counter()
            e2e_step()
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        e2e_s = statistics.median(ts)                   # the host link of a shared box hiccups: median, with the spread beside it
        e2e_spread = {"mean_ms": 1e3 * statistics.mean(ts), "min_ms": 1e3 * min(ts), "max_ms": 1e3 * max(ts), "steps": len(ts)}
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        if dist_on:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
        e2e = {"value": n_local * world / e2e_s, "unit": "nodes/s", "ms_per_step": 1e3 * e2e_s,
               "h2d_bytes_per_step": (x_host.numel() * 4 + ei_host.numel() * 8) * world, "d2h_bytes_per_step": y_host.numel() * 4 * world,
               "includes": "H2D x + edge_index (pinned), CSR build" + (", halo plan" if dist_on else "") + ", forward, D2H logits (pinned)",
               "stat": "median over steps, max over ranks", "spread_rank0": e2e_spread}

    if rank != 0:
        if dist_on:
            dist.destroy_process_group()
        return

    ms_step = ms_total / args.steps
    value = n_local * world / (ms_step * 1e-3)
    # dominant kernel: layer 0 = gather(128) -> KAN 128->64->64 -> BN ; bytes per SURVEY.md 8(d)
    S = GRID + ORDER
    top = kernels[0] if kernels else None
    roofline = None
    if top is not None:
        li = top["launch"]
        e_agg = N_EDGES
        if li < MP_LAYERS:
            f_in = N_FEAT if li == 0 else HIDDEN
            bytes_ = layer_algorithmic_bytes(n_local, e_agg, f_in, HIDDEN, kan_params([f_in, HIDDEN, HIDDEN], S))
        else:
            width = N_FEAT + MP_LAYERS * HIDDEN
            bytes_ = 4 * width * n_local + 4 * N_CLASSES * n_local + 4 * kan_params([width, N_CLASSES], S)
        ach = bytes_ / (top["ms"] * 1e-3) / 1e9
        # tensor work of the same launch: dense flops 2*N*sum(in*out*(1+S)), issued as 3 bf16 products (hi/lo split)
        if li < MP_LAYERS:
            f_in = N_FEAT if li == 0 else HIDDEN
            dense = 2.0 * n_local * (f_in * HIDDEN + HIDDEN * HIDDEN) * (1 + S)
        else:
            dense = 2.0 * n_local * (N_FEAT + MP_LAYERS * HIDDEN) * N_CLASSES * (1 + S)
        tf = 3.0 * dense / (top["ms"] * 1e-3) / 1e12
        roofline = {"bound": "hbm", "kernel": top["label"], "launch_ms": top["ms"], "algorithmic_bytes": bytes_,
                    "achieved": ach, "peak": pk["hbm_gbs"], "peak_source": pk_src, "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                    "traffic": NCU_TRAFFIC_BYTES.get(top["label"]) if not dist_on else None,
                    "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum of the same launch at N=1 (profiles/, not re-measured "
                                      "in this run); null for N>1, where the gather also reads peer memory over NVLink",
                    "tensor": {"achieved_tflops_bf16x3": tf, "peak_tflops": pk["bf16_tflops"], "frac": tf / pk["bf16_tflops"],
                               "dense_fp32_equiv_flops": dense},
                    "note": "no-reuse gather model (SURVEY 8d); feature matrix (87 MB) is L2-resident after first touch, so DRAM "
                            "traffic < algorithmic bytes; contraction on tcgen05 (bf16 hi/lo x3, fp32 accumulate in TMEM); "
                            "traffic = ncu dram bytes of the same launch (profiles/, null if that launch was not captured)"}

    cpu = None
    if not args.no_cpu_baseline and world == 1:             # the CPU baseline is timed at N = 1 only
        n_s, e_s, ts, threads = cpu_baseline_sample(sd_cpu, 1, steps=2, warmup=1)
        cpu = {"value": n_s / min(ts), "unit": "nodes/s", "cores": threads, "kind": "port",
               "sample": f"oracle forward of the full model on the full {n_s}-node / {e_s}-edge workload graph, best of 2 (1 warm-up)"}

    # ---- NVLink side of the sharded run: bytes this rank pulls from its peers per step against the measured peer-copy rate ----
    if roofline is not None and dist_on:
        widths = N_FEAT + (MP_LAYERS - 1) * HIDDEN                       # row widths gathered by the three GIN layers: 128 + 64 + 64
        if runner.mode in ("pull", "pull_overlap"):
            rows, what = int(plan.n_halo), "distinct remote rows (pulled once per layer)"
        elif runner.mode == "push":
            rows, what = int(plan.need.sum()), ("distinct remote rows: x pulled before layer 0, hidden rows pushed by their producer layer "
                                            "(masked to the ranks that reference them)")
        elif runner.mode == "peer":
            col = plan.graph.col.long()
            rows = int(((col < rank * n_local) | (col >= (rank + 1) * n_local)).sum())
            what = "every referenced remote row (in-kernel gather, no de-duplication)"
        else:
            rows, what = int(plan.n_halo), "distinct remote rows (NCCL all-to-all per layer)"
        nv_bytes = rows * widths * 4
        roofline["nvlink"] = {"bytes_per_step_per_gpu": nv_bytes, "rows": rows, "what": what,
                              "achieved_gbs_over_step": nv_bytes / (ms_step * 1e-3) / 1e9, "peak_gbs": 770.0,
                              "peak_source": "measured peer copy per direction per GPU (B200_PROFILING.md)",
                              "frac_over_step": nv_bytes / (ms_step * 1e-3) / 1e9 / 770.0,
                              "note": "rank 0's ingress; achieved = bytes / whole step time, so 1.0 would mean the step is nothing but NVLink transfer"}

    # ---- the other BASELINE configurations on this GPU (secondary numbers; N = 1 only) ----------------------------------
    extra = None
    if not dist_on and not args.no_extras:
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        import bench_extras as BX
        extra = {}
        del x_dev, ei_dev
        torch.cuda.empty_cache()
        for name in ("cora", "zinc", "mutag", "rmat"):
            try:
                fn = getattr(BX, name)
                extra[name] = fn(dev, flush, hbm_gbs=pk["hbm_gbs"]) if name == "rmat" else fn(dev, flush, with_cpu=not args.no_cpu_baseline)
            except Exception as exc:  # a secondary number must never cost the headline
                extra[name] = {"error": repr(exc)[:300]}
            torch.cuda.empty_cache()

    line = {
        "metric": METRIC, "value": value, "unit": "nodes/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "nodes_per_gpu": n_local, "edges_per_gpu": N_EDGES, "features": N_FEAT,
                   "l2": "flushed between iterations (512 MB memset)", "csr": "cached across steps (static graph)",
                   "pre_warm": ("%d untimed steps" % PRE_WARM_STEPS if dist_on else "%.2f s of untimed steps" % PRE_WARM_S) + " before the warm-up steps (clock ramp of an idle GPU)",
                   "parallelism": (f"node-range shards x{world}, " + ("remote rows gathered in-kernel over NVLink (symmetric memory), "
                                   "one device barrier per layer" if runner.mode == "peer" else ("distinct remote rows pulled over NVLink from symmetric memory "
                                   "by one copy kernel per layer (no collective)" if runner.mode == "pull" else ("x halo pulled over NVLink before layer 0; hidden rows pushed into the peers' replicas by a relay warp of the producing layer's kernel (bulk copies, masked per row), one device barrier per layer" if runner.mode == "push" else ("distinct remote rows pulled over NVLink by a copy kernel that runs concurrently with the layer (first-use order, progress counters)" if runner.mode == "pull_overlap" else "one NCCL halo all-to-all per layer")))))
                   if dist_on else "single GPU",
                   "graph": "uniform random edges over all shards: (N-1)/N of the edges are cut" if dist_on else "uniform random edges",
                   "x_halo": ("resident between steps (static features; e2e re-fetches it every step)" if args.resident_x_halo
                              else "fetched every step") if dist_on else None},
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "kernels": kernels, "roofline": roofline, "cpu_baseline": cpu,
        "extra": extra,
    }
    print(json.dumps(line), flush=True)
    if dist_on:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
