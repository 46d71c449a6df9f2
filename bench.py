#!/usr/bin/env python
"""bench.py — corrected bases/s of the features -> inference -> consensus hot path.

  python bench.py --gpus 1 --steps K --warmup W            (N>1: launched by torch.distributed.run)
  python bench.py --impl reference ...                      (the reference algorithm on the host cores)

Workload (BASELINE.json configs[1], "cfg2"): synthetic 10k reads x 15 kb, R10 error profile,
~40x coverage, `-b 64`, W = 4096.  One *step* = one device launch over `--launch-targets`
consecutive target reads of that set (cross-read batching).  With N GPUs every rank owns one
read cluster of the same size (the reference's `-c cluster` sharding): no
collective on the data path, `scaling: weak`.

Reported on ONE JSON line:
  value      corrected bases/s, whole job, device stages only, inputs already resident in HBM
             (hb_replay_last_launch; CUDA events on the launch stream)
  e2e        the same metric through the public C ABI with host buffers: hb_submit_target x T ->
             hb_flush -> hb_poll_corrected, H2D/D2H inside the timed region (host windowing and the
             one-off read-store upload are the Rust host's job and are reported separately)
  roofline   the dominant kernel class, timed live with CUDA events inside the library
  cpu_baseline  the CPU oracle (a port: the reference is Rust and cannot be built here) + torch
             fp32 forward on a bounded sample of the same targets
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "corrected_bases_per_sec"
UNIT = "bases/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=10000)
    ap.add_argument("--read-len", type=int, default=15000)
    ap.add_argument("--profile", default="r10")
    ap.add_argument("--window", type=int, default=4096)
    ap.add_argument("--batch-size", type=int, default=64, help="reference -b")
    ap.add_argument("--launch-targets", type=int, default=500)
    ap.add_argument("--cpu-sample", type=int, default=16, help="targets in the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--e2e-launch-targets", type=int, default=0, help="launch_targets of the context in the e2e region (default: same as --launch-targets; the library divides it among the submitting threads)")
    ap.add_argument("--feature-threads", type=int, default=4, help="reference -t: host threads submitting targets")
    ap.add_argument("--cpu-threads", type=int, default=0, help="threads of the CPU legs (cpu_baseline / --impl reference); 0 = all host threads")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """SM clocks / throttle reasons of the job's GPUs DURING the timed region (B200_PROFILING.md).  One sampler for the whole
    job (rank 0), through NVML in-process: a poller per rank spawning nvidia-smi five times a second contends for the driver
    with the very launches being timed (measured at N=4).  Falls back to nvidia-smi when pynvml is unavailable."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, indices):
        super().__init__(daemon=True)
        self.indices = list(indices)
        self.stop_flag = threading.Event()
        self.sm, self.mx, self.reasons = [], [], set()
        self.n = 0

    def _nvml_loop(self):
        import pynvml
        pynvml.nvmlInit()
        hs = [pynvml.nvmlDeviceGetHandleByIndex(i) for i in self.indices]
        while not self.stop_flag.is_set():
            for h in hs:
                self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                self.mx.append(float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)))
                r = int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                self.reasons |= {name for bit, name in self.REASONS.items() if r & bit}
            self.n += 1
            self.stop_flag.wait(0.1)

    def _smi_loop(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        ids = ",".join(str(i) for i in self.indices)
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={ids}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                for line in out.splitlines():
                    r = [x.strip() for x in line.split(",")]
                    if len(r) >= 6 and r[0].replace(".", "").isdigit():
                        self.sm.append(float(r[0])); self.mx.append(float(r[1]))
                        self.reasons |= {names[i] for i in range(4) if r[2 + i].lower().startswith("active")}
                self.n += 1
            except Exception:
                pass
            self.stop_flag.wait(0.5)

    def run(self):
        try:
            self._nvml_loop()
        except Exception:
            self._smi_loop()

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_min_mhz": min(self.sm) if self.sm else None,
                "sm_max_mhz": max(self.mx) if self.mx else None, "reasons": sorted(self.reasons), "samples": self.n,
                "gpus": self.indices}


def make_readset(args, rank):
    from tools import synth
    t0 = time.time()
    rs = synth.generate(args.reads, args.read_len, profile=args.profile, seed=args.seed + 1000 * rank, coverage=40.0,
                        min_ovl=2048)
    return rs, time.time() - t0


_CACHE = {}
_TUNED = {}  # host threads -> torch intra-op threads chosen by the probe in cpu_reference_run


def cpu_reference_run(rs, model, targets, window, batch_size, threads):
    """The reference algorithm on the host: C++ oracle (features, collate, consensus) on `threads`
    workers + torch fp32 forward exactly as src/inference.rs:147-175 would call the model on CPU."""
    import torch
    from concurrent.futures import ThreadPoolExecutor
    from oracle import pyoracle as po, forward_ref
    from herro_b200 import weights as hbw
    if _CACHE.get("rs") is not rs:  # the oracle's read store and the torch module are built once per read set / model
        _CACHE.update(rs=rs, reads=po.Reads(rs.ids, [rs.seq(i) for i in range(rs.n)], [rs.qual(i) for i in range(rs.n)]))
    if _CACHE.get("model") != model:
        cfg, tensors = hbw.load_blob(model)
        _CACHE.update(model=model, net=forward_ref.from_weights(cfg, tensors))
    reads, net = _CACHE["reads"], _CACHE["net"]

    def feat(t):
        ovl, cigs = rs.target_alns(t)
        return t, (po.Target(reads, t, ovl, cigs, window, batch_size) if len(ovl) else None)

    # Intra-op threads of the torch forward: "all host threads" is not the fastest setting on a many-core box (the
    # per-batch tensors are small), so the CPU arm gets the best of a few settings, measured on one batch outside the
    # timed region.  The feature/consensus legs always use `threads` workers.
    fwd_threads = _TUNED.get(threads, threads)
    if threads > 16 and threads not in _TUNED:
        probe = next((tg for _, tg in map(feat, targets[:4]) if tg is not None and tg.n_batches), None)
        if probe is not None:
            B = probe.batch(0)
            best = None
            for nt in sorted({threads, 64, 32, 16} & set(range(1, threads + 1)), reverse=True):
                torch.set_num_threads(nt)
                forward_ref.run_batch(net, B.bases, B.quals, B.lens, B.indices)  # warm
                t = time.time()
                forward_ref.run_batch(net, B.bases, B.quals, B.lens, B.indices)
                t = time.time() - t
                if best is None or t < best[0]:
                    best = (t, nt)
            fwd_threads = best[1]
        _TUNED[threads] = fwd_threads
    torch.set_num_threads(fwd_threads)
    t0 = time.time()

    with ThreadPoolExecutor(threads) as ex:
        T = list(ex.map(feat, targets))
    t_feat = time.time() - t0
    t1 = time.time()
    for _, tg in T:
        if tg is None:
            continue
        for b in range(tg.n_batches):
            B = tg.batch(b)
            info, bl = forward_ref.run_batch(net, B.bases, B.quals, B.lens, B.indices)
            for k, wi in enumerate(B.win_index):
                tg.set_logits(int(wi), info[k], bl[k])
    t_fwd = time.time() - t1
    t2 = time.time()
    bases = 0
    segs = {}
    for t, tg in T:
        if tg is None:
            continue
        s = tg.consensus()
        segs[t] = s
        bases += sum(len(x) for x in (s or []))
    t_cons = time.time() - t2
    return dict(bases=bases, seconds=time.time() - t0, t_features=t_feat, t_forward=t_fwd, t_consensus=t_cons, segments=segs,
                torch_threads=fwd_threads)


def ensure_model():
    from herro_b200 import weights as hbw
    d = os.path.join(ROOT, "tests", "_tmp")
    os.makedirs(d, exist_ok=True)
    p = os.path.join(d, f"bench_model_{os.getpid()}.hbw")
    cfg = hbw.NetConfig()
    hbw.save_blob(p, cfg, hbw.random_weights(cfg, seed=7))
    return p, cfg


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    threads = args.cpu_threads or os.cpu_count() or 1
    workload = f"cfg2: synthetic {args.reads} reads x {args.read_len} bp, {args.profile} profile, 40x, W={args.window}, -b {args.batch_size}"

    # ------------------------------------------------------------------ reference arm (CPU only)
    if args.impl == "reference":
        if rank != 0:
            return
        model, cfg = ensure_model()
        rs, _ = make_readset(args, 0)
        # bounded sample: the torch fp32 forward over whole [B,L,31] batches costs ~4 s per target read on the host,
        # so a step is 2 target reads (1 when many steps are requested) and the whole run stays within a few minutes
        per = 2 if (args.steps + args.warmup) <= 14 else 1
        need = (args.steps + args.warmup) * per
        tg = [t for t in range(rs.n)][:need]
        times, bases = [], 0
        for s in range(args.warmup + args.steps):
            r = cpu_reference_run(rs, model, tg[s * per:(s + 1) * per], args.window, args.batch_size, threads)
            if s >= args.warmup:
                times.append(r["seconds"])
                bases += r["bases"]
        tot = sum(times)
        v = bases / tot
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8+f32", "data": "synthetic",
            "config": {"workload": workload, "sample": f"{per} target reads per step"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{args.steps} steps x {per} target reads of the workload (CPU oracle on {threads} threads + torch fp32 "
                                       f"forward on {r['torch_threads']} intra-op threads, the fastest of a probe)"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        os.remove(model)
        return

    # ------------------------------------------------------------------ our arm
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (herro_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    from herro_b200 import Context
    from herro_b200 import api
    model, cfg = ensure_model()
    rs, t_gen = make_readset(args, rank)
    n_steps = args.warmup + args.steps
    lt = min(args.launch_targets, max(1, rs.n // n_steps))
    nthr = max(1, min(args.feature_threads, threads))
    lt_thread = args.e2e_launch_targets or lt   # shared by the feature threads: each stages lt / threads targets per launch (ctx.cu)
    ctx = Context(model, local_rank, args.window, args.batch_size, launch_targets=lt_thread)
    t0 = time.time()
    ctx.upload_reads(rs.seqs, rs.quals, rs.off)
    torch.cuda.synchronize()
    t_upload = time.time() - t0
    # The C++ host harness plays the Rust binary: `-t` feature threads submit targets through the C ABI,
    # one consumer thread polls corrected reads (herro_b200/host/harness.cpp).
    harness = api.HostHarness(ctx, rs.ovl9, rs.cigars, rs.cig_off, rs.aln_off, np.diff(rs.off).astype(np.uint32))
    # ---- host side that stays in the Rust binary: windowing (timed, outside the measured region)
    t0 = time.time()
    win_warm = harness.windowing(0, args.warmup * lt, nthr)
    win_timed = harness.windowing(args.warmup * lt, n_steps * lt, nthr)
    t_windowing = time.time() - t0

    harness.run(0, args.warmup * lt, nthr, win_warm)
    ctx.replay_last_launch(1)
    ctx.reset_stats()
    sampler = ClockSampler(range(args.gpus) if rank == 0 else [])
    if rank == 0:
        sampler.start()
    # ---- region 1: end to end through the C ABI, host buffers, copies inside: hb_submit_target from the
    #      feature threads (batch i+1 is staged while batch i is on the GPU), hb_poll_corrected from the consumer.
    barrier()
    t0 = time.perf_counter()
    r_e2e = harness.run(args.warmup * lt, n_steps * lt, nthr, win_timed)
    barrier()
    t_e2e = time.perf_counter() - t0
    bases_e2e = r_e2e["bases"]
    st = ctx.stats()
    # one full-size launch (lt targets), alone on the GPU: its per-kernel CUDA-event times feed the roofline
    # (in the pipelined region two lanes overlap, so per-kernel times there include the other lane's kernels),
    # and it is the launch the HBM-resident replay re-runs
    ctx.reset_stats()
    ctx.set_launch_targets(lt)
    ctx.set_kernel_timing(True)
    for t in range((n_steps - 1) * lt, n_steps * lt):
        k = t - args.warmup * lt
        ctx.submit_target(t, (int(rs.off[t + 1] - rs.off[t]) + args.window - 1) // args.window,
                          harness.ovl[int(rs.aln_off[t]):int(rs.aln_off[t + 1])],
                          win_timed[0][int(win_timed[1][k]):int(win_timed[1][k + 1])])
    ctx.flush()
    ctx.drain()
    st_full = ctx.stats()
    # ---- region 2: device stages only, inputs resident in HBM (one launch's working set is several
    #      hundred MB of matrices + activations, larger than the 126 MB L2, so no L2 flush is needed)
    last_launch_bases = st_full["last_launch_bases"]  # what the replay re-runs (a full lt-target launch)
    barrier()
    ms_dev = ctx.replay_last_launch(args.steps)
    barrier()
    sampler.stop_flag.set()
    if rank == 0:
        sampler.join(timeout=3)
    st2 = ctx.stats()
    # the replay re-runs the LAST timed launch `steps` times; its output size is known from region 1
    per_step_bases = last_launch_bases
    t_dev = ms_dev / 1e3
    vals = torch.tensor([t_dev, t_e2e, float(per_step_bases * args.steps), float(bases_e2e)], dtype=torch.float64, device="cuda")
    if dist is not None:
        tmax = vals.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = vals.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        gathered = [torch.zeros_like(vals) for _ in range(args.gpus)]
        dist.all_gather(gathered, vals)
        per_rank = {"ms_per_step": [round(1e3 * float(g[0]) / args.steps, 3) for g in gathered],
                    "bases_per_step": [float(g[2]) / args.steps for g in gathered],
                    "e2e_seconds": [round(float(g[1]), 4) for g in gathered]}
        t_dev, t_e2e = float(tmax[0]), float(tmax[1])
        bases_dev, bases_e2e_all = float(tsum[2]), float(tsum[3])
    else:
        per_rank = None
        bases_dev, bases_e2e_all = float(vals[2]), float(vals[3])

    parity_failed = False
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        tf_peak = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1590.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
        mk, nk, cf = st_full["ms_kernel"], st_full["n_kernel"], st_full["class_flops"]
        # the dominant kernel of the step, by CUDA-event time of the isolated full-size launch
        KERNEL_OF = {"gemm": "k_gemm_ws", "ffn": "k_ffn_ws", "qkv_attn": "k_qkv_attn_ws", "stem": "k_stem_tc", "pileup": "k_pass2b"}
        DESCR = {"gemm": "tcgen05 bf16x3 contractions: out-proj(+LN) and read-axis collapse",
                 "ffn": "fused FFN1 -> ReLU -> FFN2 + residual + LayerNorm on tcgen05, bf16x3",
                 "qkv_attn": "fused QKV projection (tcgen05) + read-axis attention (mma.sync), bf16x3",
                 "stem": "embedding+conv stem as a tcgen05 contraction (2 passes) + first LayerNorm",
                 "pileup": "pileup build (tile in shared memory, second get_supported, majority vote)"}
        top = max((k for k in mk if k in KERNEL_OF), key=lambda k: mk[k])
        pile_gbs = st_full["pileup_algo_bytes"] / (mk["pileup"] * 1e-3) / 1e9 if mk["pileup"] > 0 else 0.0
        traffic = {}
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r01d_traffic.json")))
        except Exception:
            pass
        t_top = traffic.get(KERNEL_OF[top], {}).get("dram_bytes_per_launch")
        if top == "pileup":
            roof = {"kernel": "k_pass2b (" + DESCR[top] + ")", "bound": "hbm", "achieved": pile_gbs, "peak": hbm_peak,
                    "unit": "GB/s", "frac": pile_gbs / hbm_peak, "traffic": t_top, "peak_source": peak_src,
                    "launches_per_step": nk[top], "ms_per_launch": mk[top] / max(nk[top], 1)}
        else:
            tf = cf[top] / (mk[top] * 1e-3) / 1e12
            roof = {"kernel": KERNEL_OF[top] + " (" + DESCR[top] + "; algorithmic fp32 FLOPs over 31 read tokens per position, "
                    "each executed as 3 bf16 MMA passes)", "bound": "tensor", "achieved": tf, "peak": tf_peak, "unit": "TFLOP/s",
                    "frac": tf / tf_peak, "traffic": t_top, "traffic_note": "avg DRAM bytes per launch, profiles/r01d_traffic.json",
                    "peak_source": peak_src, "launches_per_step": nk[top], "ms_per_launch": mk[top] / max(nk[top], 1)}
        tensor_classes = {k: {"ms": mk[k], "launches": nk[k], "algorithmic_tflops": cf[k] / (mk[k] * 1e-3) / 1e12}
                          for k in ("stem", "qkv_attn", "gemm", "ffn") if nk.get(k) and mk[k] > 0}
        out = {
            "metric": METRIC, "value": bases_dev / t_dev, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8 pileup/consensus + f32 forward", "data": "synthetic",
            "config": {"workload": workload, "targets_per_step": lt, "e2e_targets_per_launch": lt_thread, "windows_per_step": st_full["windows"],
                       "supported_positions_per_step": st_full["supported"],
                       "sharding": "one read cluster per GPU (read-id shard), no collective" if args.gpus > 1 else "single GPU",
                       "l2": "inputs larger than L2 (per-launch working set >> 126 MB)",
                       "host_feature_threads": nthr, "host_worker_busy_ms_per_launch": st["ms_worker_busy"] / max(st["device_launches"], 1),
                       "host_worker_gpu_wait_ms_per_launch": st["ms_worker_gpu_wait"] / max(st["device_launches"], 1),
                       "launches_in_e2e_region": st["device_launches"], "host_allocs_in_e2e_region": st["host_allocs"], "host_alloc_ms_in_e2e_region": st["ms_host_alloc"],
                       "submit_backpressure_ms_sum": st["ms_submit_wait"], "worker_phase_ms_per_launch": [round(x / max(st["device_launches"], 1), 3) for x in st["ms_worker_phase"][:7]], "harness_seconds": r_e2e["seconds"], "submit_seconds_sum": r_e2e["submit_seconds_sum"], "last_launch_targets": st["last_launch_targets"], "host_windowing_s": t_windowing, "read_store_upload_s": t_upload, "generate_s": t_gen,
                       "model": {"channels": cfg.channels, "heads": cfg.heads, "layers": cfg.layers, "ffn": cfg.ffn,
                                 "stem_k": cfg.stem_k, "collapse": cfg.collapse, "weights": "random init (no checkpoint offline)"}},
            "e2e": {"value": bases_e2e_all / t_e2e, "unit": UNIT,
                    "h2d_bytes_per_step": st["h2d_bytes"] / args.steps, "d2h_bytes_per_step": st["d2h_bytes"] / args.steps},
            "gpu_launches": int(st["kernel_launches"] + st2["kernel_launches"]),
            "roofline": roof,
            "kernels_ms_per_step": {k: mk[k] for k in mk if nk[k]},
            "tensor_kernels": tensor_classes,
            "pileup_roofline": {"bound": "hbm", "achieved": pile_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": pile_gbs / hbm_peak},
            "clocks": sampler.summary(),
            "per_rank": per_rank,  # each rank corrects its own read cluster: `value` = sum of bases / slowest rank's time
        }
        if not args.no_cpu_baseline and args.gpus == 1:
            tg = list(range(args.warmup * lt, args.warmup * lt + args.cpu_sample))
            r = cpu_reference_run(rs, model, tg, args.window, args.batch_size, threads)
            # parity on the measured workload: the same targets through the CUDA path, segment for segment
            ctx.set_kernel_timing(False)
            for t in tg:
                if rs.aln_off[t + 1] > rs.aln_off[t]:
                    ctx.submit_alignments(t, harness.ovl[int(rs.aln_off[t]):int(rs.aln_off[t + 1])])
            ctx.flush()
            gpu_seg = {c.rid: (c.segments or None) for c in ctx.drain()}
            same = all(gpu_seg.get(t) == r["segments"].get(t) for t in set(gpu_seg) | set(r["segments"]))
            out["parity_sample"] = {"targets": len(tg), "identical": bool(same),
                                    "bases": int(sum(len(x) for v in r["segments"].values() for x in (v or [])))}
            parity_failed = not same
            out["cpu_baseline"] = {"value": r["bases"] / r["seconds"], "unit": UNIT, "cores": threads, "kind": "port",
                                   "sample": f"{len(tg)} target reads of the workload; features {r['t_features']:.1f}s, "
                                             f"forward {r['t_forward']:.1f}s ({r['torch_threads']} torch threads, fastest of a probe), "
                                             f"consensus {r['t_consensus']:.2f}s"}
        print(json.dumps(out))
    ctx.close()
    if rank == 0 and parity_failed:
        raise SystemExit("parity_sample: the CUDA path and the CPU oracle disagree on the sampled targets")
    try:
        os.remove(model)
    except OSError:
        pass
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
