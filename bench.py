#!/usr/bin/env python
"""bench.py -- G+D training images/s of the BigGAN step (BASELINE.json metric) on N B200s.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun, one rank/GPU)
  python bench.py --impl reference ...                      (CPU restatement of the reference, host cores)

A "step" is one full training iteration: D update then G update (BigGAN.py:1061-1084) of BigGAN-128,
ch=96, global batch 256, hinge loss, n_labels=1000 (BASELINE config 1) on synthetic data of that
shape with random-init weights.  `value` = global images / step time with inputs resident in HBM;
`e2e` = the same through Trainer.train_step with the step's inputs copied from pinned host memory
and the two losses read back every step.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "G+D train img/s, BigGAN-128 b256"
UNIT = "img/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_step_rate(config_id, batch, steps, warmup, threads):
    """Times the CPU restatement of the reference (oracle, float32, torch CPU) on a bounded sample of the
    workload: the same model/config at a reduced batch.  Returns (img/s, seconds per step)."""
    import numpy as np
    import torch
    from oracle import ref_model as M, ref_ops as R
    torch.set_num_threads(threads)
    cfg = M.Config(**M.CONFIGS[config_id])
    cfg.batch_size = batch
    b = M.synthetic_batch(cfg, B=batch)
    store = R.RefStore(torch.float32)
    od = M.AdamState(cfg.d_lr, cfg.beta1, cfg.beta2)
    og = M.AdamState(cfg.g_lr, cfg.beta1, cfg.beta2, cfg.moving_decay)
    rng = np.random.Generator(np.random.PCG64(5))

    def rnd():
        return R.rnd_to_torch(R.draw_diffaug_randoms(rng, batch, cfg.img_size, cfg.img_size)) if cfg.da_policy else None
    ts = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        M.d_update(store, cfg, b["real"], b["labels"], b["z_d"], b["labels"], rnd_real=rnd(), rnd_fake=rnd(), opt_state=od)
        M.g_update(store, cfg, b["z_g"], b["labels"], rnd_fake=rnd(), opt_state=og)
        if i >= warmup:
            ts.append(time.perf_counter() - t0)
    sec = sum(ts) / len(ts)
    return batch / sec, sec


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    for var in ("OMP_NUM_THREADS", "MKL_NUM_THREADS"):       # torchrun pins these to 1; the CPU arm uses every host core
        os.environ[var] = str(threads)
    batch = args.cpu_batch
    steps, warm = max(1, min(args.steps, 3)), max(0, min(args.warmup, 1))
    t0 = time.perf_counter()
    v, sec = cpu_reference_step_rate(args.config, batch, steps, warm, threads)
    sample = "config %d model at batch %d (%d timed + %d warm-up iterations, %.1f s/iteration)" % (args.config, batch, steps, warm, sec)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args, 1),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                             "note": "torch-CPU float32 restatement of the reference (TensorFlow 1.x is not installable here)"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.perf_counter() - t0}
    print(json.dumps(line))


def workload_config(args, world):
    from oracle.ref_model import CONFIGS
    c = CONFIGS[args.config]
    return {"workload": "BigGAN-%d G+D training iteration (D update + G update), ch=%d, global batch %d, %d classes, hinge, "
                        "da_policy='%s' (BASELINE config %d)" % (c["img_size"], c["ch"], args.batch or c["batch_size"], c["n_labels"],
                                                                 c.get("da_policy", ""), args.config),
            "global_batch": args.batch or c["batch_size"], "per_gpu_batch": (args.batch or c["batch_size"]) // world,
            "parallelism": "dp%d" % world + ("" if world == 1 else " (%s, %s, BN statistics %s)" % (
                "all-reduce + replicated Adam" if args.no_shard_optimizer else "reduce-scatter + sharded Adam + all-gather",
                "no overlap" if args.no_overlap else "D update overlapped with the G-phase generator forward",
                "over NCCL" if os.environ.get("BG_BN_SYNC", "p2p") != "p2p" else "over NVLink peer memory")),
            "l2": "no explicit flush: every step streams several GB of activations/weights, far larger than the 126 MB L2"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=1)
    ap.add_argument("--batch", type=int, default=0, help="override the global batch (debug)")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--cpu-batch", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-json", default="", help="write the per-entry-point timing breakdown here")
    ap.add_argument("--no-graph", action="store_true", help="run the eager step instead of the captured CUDA graph")
    ap.add_argument("--no-shard-optimizer", action="store_true",
                    help="N > 1: all-reduce + replicated Adam instead of reduce-scatter, Adam on 1/N of the arena per rank, all-gather")
    ap.add_argument("--no-overlap", action="store_true", help="N > 1: keep the D update's collective + Adam on the compute stream")
    ap.add_argument("--quick", action="store_true", help="warm-up + timed steps only (for ncu captures): no e2e / breakdown pass")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local)
    pg = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        pg = dist.group.WORLD

    import __graft_entry__
    if not os.path.exists(os.path.join(ROOT, "biggan_tensorflow_b200", "libbiggan_b200.so")):
        if local == 0:
            __graft_entry__.build()
        if world > 1:
            dist.barrier()
    from biggan_tensorflow_b200 import _lib
    from biggan_tensorflow_b200.model import CONFIGS, Config
    from biggan_tensorflow_b200.step import Trainer
    from oracle.ref_model import synthetic_batch, Config as RConfig

    kw = dict(CONFIGS[args.config])
    if args.batch:
        kw["batch_size"] = args.batch
    cfg = Config(**kw)
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    tr = Trainer(cfg, compute_dtype=dtype, rank=rank, world_size=world, process_group=pg, shard_optimizer=not args.no_shard_optimizer, overlap=not args.no_overlap)
    tr.build()
    B, Bl = cfg.batch_size, cfg.batch_size // world
    sb = synthetic_batch(RConfig(**kw), B=B)           # same synthetic tensors on every rank; each takes its shard
    sl = slice(rank * Bl, (rank + 1) * Bl)
    host = {k: torch.from_numpy(np.ascontiguousarray(v[sl])).pin_memory() for k, v in sb.items()}
    devt = {k: v.cuda(non_blocking=True) for k, v in host.items()}
    h2d = sum(v.numel() * v.element_size() for v in host.values())

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    graphed = None
    if not args.no_graph and not cfg.da_policy:
        from biggan_tensorflow_b200.graph import GraphedStep
        try:
            graphed = GraphedStep(tr, devt["real"], devt["labels"], devt["z_d"], devt["z_g"], devt["labels"])
        except Exception as e:           # capture is an optimisation: the eager CUDA step is the same computation
            print("[bench] CUDA graph capture failed, running eagerly: %r" % (e,), file=sys.stderr)
            graphed = None
    run = graphed if graphed is not None else tr.train_step

    def step_resident():
        return run(devt["real"], devt["labels"], devt["z_d"], devt["z_g"], devt["labels"])

    def step_e2e():
        if graphed is not None:          # H2D straight into the graph's static input buffers
            dl, gl = run(host["real"], host["labels"], host["z_d"], host["z_g"], host["labels"])
        else:
            d = {k: v.cuda(non_blocking=True) for k, v in host.items()}
            dl, gl = run(d["real"], d["labels"], d["z_d"], d["z_g"], d["labels"])
        return float(dl.cpu()), float(gl.cpu())        # device->host read of the step's result

    def timed(fn, steps):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)      # max over ranks, device-timed
        return float(ms.cpu()) / steps, out

    def finish():
        """All ranks: drain, drop the captured graph, tear NCCL down, and leave without running interpreter
        teardown (a captured graph holding NCCL work can otherwise stall process exit)."""
        nonlocal graphed
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        if world > 1:
            # every collective of this rank has completed (synchronize above).  Leave without NCCL / CUDA-graph
            # teardown: destroying a communicator that a live captured graph references can stall process exit.
            os._exit(0)

    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count
    ms_step, _ = timed(step_resident, args.steps)
    launches = graphed.launches if graphed is not None else (_lib.launch_count - l0) // args.steps
    clocks = sampler.stop() if rank == 0 else None
    if args.quick:
        if rank == 0:
            print(json.dumps({"quick": True, "ms_per_step": ms_step, "value": B / (ms_step * 1e-3), "gpu_launches": launches}))
        return finish()
    step_e2e()
    ms_e2e, losses = timed(step_e2e, max(2, args.steps // 2))

    # per-entry-point breakdown of one more step (CUDA events around every C-ABI call on the launch stream)
    _lib.profile_begin()
    tr.train_step(devt["real"], devt["labels"], devt["z_d"], devt["z_g"], devt["labels"])      # eager, so that events bracket each call
    prof = _lib.profile_end()
    if rank != 0:
        return finish()
    peaks = load_peaks()
    tot_ms = sum(d["ms"] for d in prof.values())
    conv = {k: d for k, d in prof.items() if k.startswith("conv_") and "[fc]" not in k}
    c_ms = sum(d["ms"] for d in conv.values())
    c_fl = sum(d["flops"] for d in conv.values())
    c_calls = sum(d["calls"] for d in conv.values())
    ach = c_fl / (c_ms * 1e-3) / 1e12 if c_ms > 0 else 0.0
    peak = peaks["bf16_tflops_sustained"] if dtype == torch.bfloat16 else 75.0
    step_flops = {0: 0.87e9, 1: 212.7e9, 2: 212.7e9, 3: 714.0e9, 4: 603.6e9}[args.config] * B
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "r3_conv_traffic.json")
    if os.path.exists(tp) and args.config == 1 and dtype == torch.bfloat16 and world == 1 and not args.batch:
        t = json.load(open(tp))      # dram__bytes_read.sum + dram__bytes_write.sum of the same kernels, per launch, one ncu capture
        traffic, traffic_src = t["dram_bytes_per_launch"], "%s (ncu, %d launches of one iteration)" % (t["source"], t["launches_per_step"])
    roof = {"bound": "tensor", "kernel": "tc_fwd_kernel/tc_fwd_halo_kernel/tc_wgrad_kernel (conv, deconv, dgrad, wgrad implicit GEMMs)",
            "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": traffic, "traffic_unit": "bytes per launch (DRAM)",
            "traffic_source": traffic_src,
            "peak_source": "%s bf16_tflops_sustained (kernel timed inside a long step)" % peaks["source"],
            "launches_per_step": c_calls, "avg_launch_ms": c_ms / max(c_calls, 1), "share_of_step": c_ms / max(tot_ms, 1e-9),
            "how": "CUDA events around every conv C-ABI call of one extra step on the launch stream; achieved = sum of "
                   "algorithmic conv FLOPs / sum of those durations",
            "step_algorithmic_tflop": step_flops / 1e12,
            "whole_step_frac_of_peak": step_flops / (ms_step * 1e-3) / 1e12 / (peak * world)}
    if args.profile_json:
        os.makedirs(os.path.dirname(os.path.abspath(args.profile_json)), exist_ok=True)
        json.dump({"ms_per_step": ms_step, "profiled_step_ms_sum": tot_ms,
                   "by_entry_point": dict(sorted(prof.items(), key=lambda kv: -kv[1]["ms"]))}, open(args.profile_json, "w"), indent=1)
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        threads = os.cpu_count() or 1
        v, sec = cpu_reference_step_rate(args.config, args.cpu_batch, 1, 1, threads)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "config %d model at batch %d on the host cores (1 timed + 1 warm-up iteration, %.1f s/iteration)" % (args.config, args.cpu_batch, sec),
               "note": "torch-CPU float32 restatement of the reference; TensorFlow 1.x (the reference runtime) is not installable here"}
    line = {"metric": METRIC, "value": B / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic", "config": workload_config(args, world),
            "e2e": {"value": B / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8,
                    "ms_per_step": ms_e2e, "losses": list(losses)},
            "gpu_launches": launches, "cuda_graph": graphed is not None, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu}
    print(json.dumps(line))
    finish()


if __name__ == "__main__":
    main()
