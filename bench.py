#!/usr/bin/env python
"""Benchmark of the WeKws streaming KWS forward path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

Workload (BASELINE.json configs[1]): the `mdtc` model (hidden 64, 1+4x4 DS-dilated blocks, k=5),
80-dim features, B = 1024 concurrent streams per GPU, chunk T = 40 frames, streaming caches
(B,64,244) carried from step to step.  A "step" is one pass of KWSModel.forward over one batch of
B x T synthetic frames.  Metric: audio-hours/sec = frames/sec / 360000 (100 frames per second).

* `value`      whole-job throughput, inputs resident in HBM, device-timed (CUDA events), max over ranks.
* `e2e`        same metric through the public Python surface (wekws_b200.KWSModel.forward) with the
               step's features coming from pinned HOST memory (H2D) and its posteriors read back (D2H)
               inside the timed region; the streaming cache stays on the device as it does in the
               reference's own streaming loop (stream_kws_ctc.py:487).
* `roofline`   dominant kernel conv_backbone_kernel<64>: algorithmic bytes per launch
               (idim*4 + odim*4 + 2*cache_bytes_per_stream/T = 3447 B/frame) / mean
               launch duration measured with CUDA events inside the timed region, against the measured
               HBM copy bandwidth in MEASURED_PEAKS.json.
* `cpu_baseline` the oracle port of the reference's PyTorch CPU path (oracle/kws_oracle.py) timed on the
               host cores on a bounded sample of the same workload.
Multi-GPU: streams are independent -> each rank owns its own B streams ("weak" scaling), no collective
on the data path; torch.distributed (NCCL) is used only for the barrier and the max-over-ranks time.
Between timed steps the working set rotates over NSETS independent stream sets (> 126 MB L2).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FRAMES_PER_HOUR = 360000.0
WORKLOADS = {
    # name: (model, B per GPU, T, idim)
    "mdtc_b1024_t40": ("mdtc", 1024, 40, 80),
    "gru_b512_t1": ("gru", 512, 1, 80),
    "tcn_b1024_t40": ("tcn", 1024, 40, 80),
    "ds_tcn_b1024_t40": ("ds_tcn", 1024, 40, 80),
    # BASELINE configs[4]: raw PCM -> posterior, 10 000 one-second clips over 8 GPUs = 1250 clips per GPU;
    # T = 98 frames per clip (16000 samples), every clip starts a stream (no cache carried)
    "pcm_e2e_1250x1s": ("mdtc", 1250, 98, 80),
}
PCM_SAMPLES = 16000
NSETS = 4


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes_per_frame(model, T, idim, odim=1):
    """features in + posteriors out + streaming cache read once and written once."""
    cache_bytes = {"mdtc": 64 * 244 * 4, "mdtc_small": 32 * 184 * 4, "ds_tcn": 256 * 105 * 4,
                   "tcn": 64 * 105 * 4, "gru": 2 * 128 * 4}[model]
    return idim * 4 + odim * 4 + 2.0 * cache_bytes / T


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed regions (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def build_oracle_model(model_name, idim):
    from wekws_b200 import init_model, model_config, synth
    cfg = model_config(model_name, input_dim=idim)
    torch.manual_seed(777)
    m = synth.randomize_(init_model(cfg), seed=777).eval()
    return cfg, m


def cpu_reference_run(model_name, B, T, idim, steps, warmup, budget_s=20.0):
    """The reference's CPU PyTorch path (oracle port, same ATen ops) on all host cores, on a bounded
    sample of the workload: Bs streams x T frames per step, caches carried."""
    from oracle import kws_oracle as O
    from wekws_b200 import synth
    ncpu = os.cpu_count() or 1
    cfg, m = build_oracle_model(model_name, idim)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    Bs = min(B, 256)           # survey: best CPU throughput of this model is at B=256
    x = synth.features(Bs, T, idim, seed=4321)
    gru = cfg["backbone"]["type"] == "gru"

    def fresh_cache():
        return torch.zeros(cfg["backbone"]["num_layers"], Bs, cfg["hidden_dim"]) if gru else None

    # give the reference its best shot: probe thread counts (all cores oversubscribes ATen's small convs)
    cand = sorted({n for n in (1, 4, 8, 16, 32, 64, ncpu) if n <= ncpu})
    best, cores = None, ncpu
    for n in cand:
        torch.set_num_threads(n)
        c = fresh_cache()
        _, c = O.kws_forward(sd, cfg, x, c)
        t0 = time.perf_counter()
        _, c = O.kws_forward(sd, cfg, x, c)
        dt = time.perf_counter() - t0
        if best is None or dt < best:
            best, cores = dt, n
    torch.set_num_threads(cores)
    cache = fresh_cache()
    for _ in range(warmup):
        _, cache = O.kws_forward(sd, cfg, x, cache)
    times = []
    t_start = time.perf_counter()
    for _ in range(steps):
        t0 = time.perf_counter()
        _, cache = O.kws_forward(sd, cfg, x, cache)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > budget_s and len(times) >= 3:
            break
    total = sum(times)
    fps = Bs * T * len(times) / total
    return {"value": fps / FRAMES_PER_HOUR, "unit": "audio-hours/s", "cores": cores, "kind": "port",
            "frames_per_sec": fps, "ms_per_step": 1e3 * total / len(times), "steps": len(times),
            "sample": f"oracle port (torch CPU, best of {cand} threads = {cores}; host has {ncpu}) of {model_name}: "
                      f"{Bs} streams x {T} frames per step, cache carried, {len(times)} steps"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="mdtc_b1024_t40", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    model_name, B, T, idim = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": f"{model_name}: {B} streams/GPU x {T}-frame chunks, {idim}-dim features, cache carried",
              "model_cfg": model_name, "streams_per_gpu": B, "chunk_frames": T, "feature_dim": idim,
              "global_streams": B * max(args.gpus, 1), "parallelism": f"streams sharded x{max(args.gpus, 1)} (no collective)",
              "l2": f"working set rotates over {NSETS} independent stream sets (~{NSETS * 77} MB) > 126 MB L2"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        steps = min(args.steps, 50)
        r = cpu_reference_run(model_name, B, T, idim, steps, max(min(args.warmup, 3), 1), budget_s=60.0)
        line = {"impl": "reference", "metric": "audio-hours/sec KWS scoring (frames/sec / 360000)",
                "value": r["value"], "unit": "audio-hours/s", "n_gpus": args.gpus, "steps": r["steps"],
                "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "frames_per_sec": r["frames_per_sec"],
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": "audio-hours/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (B200)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)
    from wekws_b200 import _native, init_model, model_config, synth
    cfg = model_config(model_name, input_dim=idim)
    torch.manual_seed(777)
    model = synth.randomize_(init_model(cfg), seed=777).eval().to(dev)
    gru = model_name == "gru"
    cshape = (2, B, 128) if gru else (B, model.hdim, model.backbone.padding)
    pcm_mode = args.workload.startswith("pcm_")
    if pcm_mode:
        from wekws_b200 import Fbank
        fb = Fbank(idim)
        pcm = [synth.pcm_int16(B, PCM_SAMPLES, seed=1234 + rank * 17 + s).to(dev) for s in range(NSETS)]
        feat_buf = torch.empty(B, T, idim, device=dev)
        feats = [feat_buf] * NSETS
        caches = [None] * NSETS
        config["workload"] = (f"raw int16 PCM -> Fbank -> {model_name}: {B} one-second clips/GPU "
                              f"({T} frames each), start of stream")
    else:
        feats = [synth.features(B, T, idim, seed=4321 + rank * 17 + s).to(dev) for s in range(NSETS)]
        caches = [torch.zeros(cshape, device=dev) for _ in range(NSETS)]

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput (`value`) + per-launch durations (roofline) ----------
    def step(i):
        s = i % NSETS
        if pcm_mode:
            model(fb(pcm[s], out=feat_buf))
        else:
            _, caches[s] = model(feats[s], caches[s])

    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = _native.launch_count()
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin.record()
    for i in range(args.steps):
        step(i)
    t_end.record()
    barrier()
    launches = _native.launch_count() - n0
    total_ms = t_begin.elapsed_time(t_end)
    # per-launch durations (roofline, p50/p99) from a separate short pass: bracketing every step with
    # events costs more host time than a 20 us kernel takes, so it stays out of the timed region
    nlat = min(args.steps, 200)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(nlat)]
    for i in range(nlat):
        ev[i][0].record()
        step(i)
        ev[i][1].record()
    barrier()
    per_launch_ms = [a.elapsed_time(b) for a, b in ev]
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    frames_total = B * T * args.steps * world
    fps = frames_total / (total_ms_max * 1e-3)

    # ---------------- end to end through the public API with host buffers (`e2e`) ----------------------
    h_feats = [(p if pcm_mode else f).cpu().pin_memory() for p, f in zip(pcm if pcm_mode else feats, feats)]
    h_out = [torch.empty(B, T, model.odim).pin_memory() for _ in range(2)]
    d_in = [torch.empty_like(pcm[0] if pcm_mode else feats[0]) for _ in range(2)]
    copy_stream, main_stream = torch.cuda.Stream(dev), torch.cuda.current_stream(dev)
    in_ready = [torch.cuda.Event() for _ in range(2)]
    in_free = [torch.cuda.Event() for _ in range(2)]

    def e2e_run(nsteps):
        # H2D of step i+1 overlaps the kernel of step i (separate copy stream, double buffered);
        # every step's posteriors are copied back to pinned host memory
        with torch.cuda.stream(copy_stream):
            d_in[0].copy_(h_feats[0], non_blocking=True)
            in_ready[0].record(copy_stream)
        for i in range(nsteps):
            b, s = i & 1, i % NSETS
            if i + 1 < nsteps:
                nb = (i + 1) & 1
                with torch.cuda.stream(copy_stream):
                    if i >= 1:
                        copy_stream.wait_event(in_free[nb])
                    d_in[nb].copy_(h_feats[(i + 1) % NSETS], non_blocking=True)
                    in_ready[nb].record(copy_stream)
            main_stream.wait_event(in_ready[b])
            if pcm_mode:
                y, _ = model(fb(d_in[b], out=feat_buf))
            else:
                y, caches[s] = model(d_in[b], caches[s])
            in_free[b].record(main_stream)
            h_out[b].copy_(y, non_blocking=True)
        main_stream.synchronize()

    e2e_steps = args.steps
    e2e_run(max(args.warmup, 3))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.perf_counter()
    e0.record()
    e2e_run(e2e_steps)
    e1.record()
    barrier()
    # the sampler spans both timed loops (device-resident + end to end): a 2000-step GRU loop is over in 40 ms,
    # shorter than one nvidia-smi sampling period
    clocks = sampler.stop() if rank == 0 else None
    e2e_ms = max(e0.elapsed_time(e1), 0.0)
    e2e_wall_ms = (time.perf_counter() - wall0) * 1e3
    t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_fps = B * T * e2e_steps * world / (float(t.item()) * 1e-3)

    # ---------------- single-stream chunk latency (BASELINE configs[0] shape on the GPU) ---------------
    lat = None
    if rank == 0:
        x1 = synth.features(1, T, idim, seed=1).to(dev)
        c1 = torch.zeros((2, 1, 128) if gru else (1, model.hdim, model.backbone.padding), device=dev)
        for _ in range(20):
            _, c1 = model(x1, c1)
        torch.cuda.synchronize()
        ls = []
        for _ in range(200):
            a, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _, c1 = model(x1, c1)
            b2.record()
            torch.cuda.synchronize()
            ls.append(a.elapsed_time(b2))
        lat = statistics.median(ls)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    peak, peak_src = measured_peaks()
    bpf = algorithmic_bytes_per_frame(model_name, T, idim, model.odim)
    if pcm_mode:   # PCM in + posteriors out + the new cache written once (no cache read at start of stream)
        bpf = PCM_SAMPLES * 2.0 / T + model.odim * 4 + (64 * 244 * 4) / T
    launch_ms = statistics.mean(per_launch_ms)
    srt = sorted(per_launch_ms)
    p99_ms = srt[min(len(srt) - 1, int(0.99 * len(srt)))]
    achieved = bpf * B * T / (launch_ms * 1e-3) / 1e9
    traffic = None
    prof = os.path.join(ROOT, "profiles", "dominant_kernel.json")
    if os.path.exists(prof):
        try:
            with open(prof) as f:
                traffic = json.load(f).get(args.workload, {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    flop_per_frame = {"mdtc": 299776, "tcn": 272512, "ds_tcn": 582144, "gru": 413952}.get(model_name)
    line = {
        "metric": "audio-hours/sec KWS scoring (frames/sec / 360000)",
        "value": fps / FRAMES_PER_HOUR, "unit": "audio-hours/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
        "frames_per_sec": fps,
        "p50_step_latency_ms": statistics.median(per_launch_ms), "p99_step_latency_ms": p99_ms,
        "p50_chunk_latency_ms_1stream": lat,
        "clocks": clocks,
        "e2e": {"value": e2e_fps / FRAMES_PER_HOUR, "unit": "audio-hours/s", "frames_per_sec": e2e_fps,
                "h2d_bytes_per_step": (B * PCM_SAMPLES * 2) if pcm_mode else (B * T * idim * 4),
                "d2h_bytes_per_step": B * T * model.odim * 4,
                "ms_per_step": float(t.item()) / e2e_steps,
                "api": "wekws_b200.KWSModel.forward(feats, cache); pinned host feats in, posteriors out, "
                       "H2D double-buffered on a copy stream"},
        "gpu_launches": launches,
        "tensor_cores": bool(model.uses_tensor_cores(T)),
        "roofline": {"kernel": ("fbank_kernel + " if pcm_mode else "") + (({"tcn": "tcn_tc_kernel", "ds_tcn": "dstcn_tc_kernel"}.get(model_name, "mdtc_tc_kernel") if model.uses_tensor_cores(T) else "conv_backbone_kernel") if not gru else "gru_kernel"), "bound": "hbm",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "peak_source": peak_src + ", burst", "traffic": traffic,
                     "algorithmic_bytes_per_frame": bpf, "launch_ms": launch_ms,
                     "fp32_tflops": (flop_per_frame * B * T / (launch_ms * 1e-3) / 1e12) if flop_per_frame else None,
                     "note": "HBM fraction as SURVEY 8d asks; the fused backbone is instruction/shared-memory bound "
                             "(DESIGN.md section 4)"},
    }
    # SURVEY 8d second figure for the tensor-core kernels: bf16 tensor-pipe utilisation = dense GEMM FLOP/frame x 3
    # passes of the bf16x3 split / measured bf16 peak (MEASURED_PEAKS.json); informational, never fatal
    try:
        dense = {"mdtc": 34 * 2 * 64 * 64 + 2 * 80 * 64, "tcn": 32 * 2 * 64 * 64 + 2 * 80 * 64,
                 "ds_tcn": 4 * 2 * 256 * 256 + 2 * 80 * 256}.get(model_name)
        if dense and line["tensor_cores"] and not pcm_mode:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                pk = json.load(f)
            bf16_peak = float(pk.get("bf16_tflops", pk.get("bf16_tflops_burst", 0)) or 0)
            tf = 3.0 * dense * B * T / (launch_ms * 1e-3) / 1e12
            line["roofline"]["tensor_pipe"] = {"achieved_bf16_tflops": tf, "peak_bf16_tflops": bf16_peak or None,
                                               "frac": (tf / bf16_peak) if bf16_peak else None, "passes": 3}
    except Exception:
        pass
    if not args.no_cpu_baseline and world == 1:
        r = cpu_reference_run(model_name, B, T, idim, steps=40, warmup=2, budget_s=15.0)
        line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
