#!/usr/bin/env python
"""Headline benchmark: streamed audio-seconds per second of the 2-layer GRU-128 KWS path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): streaming serve of
S = 131,072 concurrent streams PER GPU with carried GRU state and VAD reset, 300 ms chunks of
synthetic 16 kHz int16 audio, random-init weights.  One "step" = one 300 ms chunk for every stream:
VAD -> tail carry -> framing/|rFFT|/mel -> 2x GRU(128) -> FC -> softmax -> window decode -> trigger.
Streams are independent, so N GPUs hold N*S streams with no collective on the hot path (weak scaling);
NCCL is used only for the max-over-ranks of the device time and the final trigger-count gather.

One JSON line on rank 0: `value` = whole-job audio-s/s with the chunk already resident in HBM,
`e2e` = the same through the host-buffer C-ABI call (pinned host PCM -> H2D -> step -> D2H triggers,
copies inside the timed region), `roofline` for the dominant kernel, `cpu_baseline` = the CPU oracle
port timed on this box's host cores on a bounded sample.  `--impl reference` times that CPU port alone.
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

CHUNK = 4800                      # 300 ms at 16 kHz (README.md:88, detector.py:150)
FRAMES = 30                       # steady state: 320 carried + 4800 new samples -> 30 frames
AUDIO_S_PER_CHUNK = CHUNK / 16000.0
METRIC = "streamed audio-sec/sec, 2L GRU-128 KWS"
UNIT = "audio-s/s"
# algorithmic work: GRU 325,632 + FC 1,536 FLOP per frame
GRU_FLOP_PER_FRAME = 2 * (40 * 384 + 128 * 384 + 128 * 384 + 128 * 384) + 2 * 128 * 6
# front-end kernel, compulsory bytes per stream-chunk: 4800 int16 in, carried tail 320 samples read + written,
# 30 x 40 fp32 mel out, VAD flag + frame count + tail length
FE_BYTES_PER_STREAM_CHUNK = CHUNK * 2 + 2 * 320 * 2 + FRAMES * 40 * 4 + 9


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=131072, help="concurrent streams per GPU")
    ap.add_argument("--precision", default="tc", choices=["tc", "fp32"],
                    help="recurrent kernel: tc = tcgen05 fp16 operands / fp32 accumulate (default), fp32 = exact FFMA path")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--waves", type=int, default=16, help="e2e: the streams of a GPU are served in this many waves "
                    "(independent stream objects on their own CUDA streams) so copies overlap compute and a chunk's "
                    "latency is one wave's, not the whole batch's")
    ap.add_argument("--depth", type=int, default=2, help="e2e: waves in flight")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-bind", action="store_true", help="do not pin the rank to its GPU's NUMA-local cores")
    ap.add_argument("--cpu-streams", type=int, default=0, help="CPU sample: streams (default 64 per host core)")
    ap.add_argument("--cpu-chunks", type=int, default=0, help="CPU sample: chunks per step (default: calibrated)")
    return ap.parse_args()


# ----------------------------------------------------------------------------- CPU port (oracle)
def _cpu_worker(args):
    """One process = one host core: the detector loop of the oracle for a block of streams."""
    seed, n_streams, n_chunks = args
    os.environ["OMP_NUM_THREADS"] = "1"
    try:
        from threadpoolctl import threadpool_limits
        limiter = threadpool_limits(limits=1)
    except Exception:       # pragma: no cover
        limiter = None
    import numpy as np
    from oracle import model as om, streaming as ost
    w = om.init_weights(seed=1234, n_mel=40)
    rng = np.random.default_rng(seed)
    pcm = np.clip(np.rint(rng.standard_normal((n_streams, CHUNK * (n_chunks + 1))) * 600), -32768, 32767).astype(np.int16)
    orc = ost.StreamOracle(w, n_streams)
    orc.step(pcm[:, :CHUNK])                        # warm-up chunk (fills the tail: steady state)
    t0 = time.perf_counter()
    trig = 0
    for c in range(1, n_chunks + 1):
        trig += int(orc.step(pcm[:, c * CHUNK:(c + 1) * CHUNK])["trigger"].sum())
    dt = time.perf_counter() - t0
    del limiter
    return dt, trig


def cpu_port_throughput(total_streams, n_chunks, procs=None):
    """audio-s/s of the oracle port using `procs` single-threaded processes over disjoint streams."""
    import multiprocessing as mp
    procs = procs or (os.cpu_count() or 1)
    procs = max(1, min(procs, total_streams))
    per = max(1, total_streams // procs)
    jobs = [(1000 + i, per, n_chunks) for i in range(procs)]
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(procs) as pool:
        res = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    slowest = max(r[0] for r in res)                # the processes run concurrently
    audio = per * procs * n_chunks * AUDIO_S_PER_CHUNK
    return dict(value=audio / slowest, unit=UNIT, cores=procs, kind="port",
                sample="%d streams x %d chunks of 300 ms through the numpy oracle (detector loop), %d single-thread "
                       "processes; %.1f s wall incl. start-up" % (per * procs, n_chunks, procs, wall))


def cpu_sample_size(args, budget_s):
    """Pick the bounded CPU sample: 64 streams per core (the port is fastest around there); chunk count calibrated to ~budget_s per sample."""
    cores = os.cpu_count() or 1
    streams = args.cpu_streams or 64 * cores
    if args.cpu_chunks:
        return streams, args.cpu_chunks
    probe = cpu_port_throughput(streams, 2)
    t_chunk = streams * AUDIO_S_PER_CHUNK / probe["value"]
    return streams, int(max(2, min(400, budget_s / max(t_chunk, 1e-3))))


def run_reference_arm(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path.  TensorFlow/librosa are not
    installable here, so this is the oracle port (kind 'port') on all host cores."""
    if rank != 0:
        return
    streams, chunks = cpu_sample_size(args, budget_s=150.0 / max(1, args.warmup + args.steps))
    vals = []
    info = None
    for i in range(args.warmup + args.steps):       # every step = one bounded sample of the workload
        info = cpu_port_throughput(streams, chunks)
        if i >= args.warmup:
            vals.append(info["value"])
    value = statistics.median(vals)
    info["value"] = value
    args.cpu_streams, args.cpu_chunks = streams, chunks
    per_step_audio = args.cpu_streams * args.cpu_chunks * AUDIO_S_PER_CHUNK
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=len(vals), warmup=args.warmup,
                ms_per_step=1e3 * per_step_audio / value, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32", data="synthetic", impl="reference",
                config=dict(workload="streaming serve (BASELINE configs[2]): concurrent streams with carried GRU state and "
                                     "VAD reset, 300 ms chunks; bounded CPU sample", streams_per_step=args.cpu_streams,
                            chunk_samples=CHUNK),
                cpu_baseline=info,
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, power, reasons = [], [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    power_w_max=max(power) if power else None, samples=len(sm), reasons=sorted(reasons))


# ----------------------------------------------------------------------------- ours
def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=float(d["hbm_gbs"]), bf16=float(d["bf16_tflops"]), bf16_sustained=float(d["bf16_tflops_sustained"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


def load_traffic(S):
    """Per-launch DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) from the committed ncu --set full
    capture, scaled from the captured stream count to S; None when the capture file is absent."""
    path = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if not os.path.exists(path):
        return {}
    d = json.load(open(path))
    k = S / float(d["streams"])
    out = {name: (v["dram_bytes_read"] + v["dram_bytes_write"]) * k for name, v in d["kernels"].items()}
    out["_source"] = d["source"]
    return out


def make_chunks(torch, S, n_buf, device, seed):
    """n_buf distinct [S, 4800] int16 chunks on the device: per-stream noise level log-uniform in [30, 3000] LSB,
    ~30% of (stream, chunk) cells near-silent so the VAD reset fires."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    bufs = []
    for b in range(n_buf):
        sigma = torch.exp(torch.empty((S, 1), device=device).uniform_(3.4, 8.0, generator=g))
        x = torch.randn((S, CHUNK), device=device, generator=g) * sigma
        silent = torch.rand((S, 1), device=device, generator=g) < 0.3
        x = torch.where(silent, torch.randint(-2, 3, (S, CHUNK), device=device, generator=g).float(), x)
        bufs.append(x.clamp_(-32768, 32767).round_().to(torch.int16).contiguous())
        del x
    return bufs


def timed_loop(torch, fn, steps, stream, barrier):
    """EXACTLY `steps` calls bracketed by barrier + synchronize; device time via CUDA events on `stream`."""
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    barrier()
    torch.cuda.synchronize()
    ev[0].record(stream)
    for i in range(steps):
        fn(i)
        ev[i + 1].record(stream)
    torch.cuda.synchronize()
    barrier()
    per_step = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
    return ev[0].elapsed_time(ev[steps]), per_step


def run_ours(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    barrier = (lambda: dist.barrier()) if world > 1 else (lambda: None)

    from keyword_spotting_b200 import Config, DeployModel, ModelWeights, StreamingDetector, _lib, _tensors, sharding

    host_binding = sharding.bind_host_to_gpu(local_rank) if not args.no_bind else {"bound": False, "why": "--no-bind"}
    cfg = Config(n_mel=40)
    model = DeployModel(cfg, ModelWeights.random_init(cfg, seed=1234), device=device, precision=args.precision)
    S = args.streams
    det = StreamingDetector(model, S)
    n_buf = 4                                        # 4 x 1.26 GB of PCM >> 126 MB L2: every step reads HBM-cold input
    chunks = make_chunks(torch, S, n_buf, device, seed=5678 + rank)
    stream = torch.cuda.current_stream(device)
    trig_total = torch.zeros((), dtype=torch.int64, device=device)
    trig = torch.empty(S, dtype=torch.int32, device=device)
    lib = _lib.load()

    def step_dev(i):
        x = chunks[i % n_buf]
        _lib.check(lib.kws_stream_step(det._handle, x.data_ptr(), CHUNK, x.stride(0), trig.data_ptr(), None, None,
                                       stream.cuda_stream))

    for i in range(max(args.warmup, 3)):
        step_dev(i)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                              # 50 ms samples over every timed region below
    total_ms, per_step = timed_loop(torch, step_dev, args.steps, stream, barrier)
    total_ms_max = sharding.reduce_max_scalar(total_ms, device)
    ms_per_step = total_ms_max / args.steps
    value = world * S * AUDIO_S_PER_CHUNK * args.steps / (total_ms_max * 1e-3)
    trig_total += trig.sum()

    # ---- the parts of the production step (front end incl. fused VAD/tail | GRU layers | decode), CUDA events inside
    # kws_stream_step (debug hook: every step synchronises, so this runs after the timed region)
    import ctypes
    lib.kws_debug_step_timing(1, None, None)
    for i in range(2):                               # first use of the hook: event creation, not measured
        step_dev(i)
    lib.kws_debug_step_timing(0, None, None)
    lib.kws_debug_step_timing(1, None, None)         # switching resets the sums
    for i in range(6):
        step_dev(i)
    torch.cuda.synchronize()
    part_ms = (ctypes.c_double * 3)()
    part_n = ctypes.c_longlong(0)
    lib.kws_debug_step_timing(0, part_ms, ctypes.byref(part_n))
    parts = [part_ms[i] / max(1, part_n.value) for i in range(3)]

    # ---- the same kernels alone through the public entry points (front end without the fused pre-step, GRU from
    # row-major mel) on the same data
    mel = torch.empty((S, FRAMES, cfg.n_mel), dtype=torch.float32, device=device)
    pcm_full = torch.cat([chunks[0][:, -320:], chunks[1]], dim=1).contiguous()      # a steady-state 5120-sample input
    _lib.check(lib.kws_frontend_mel(model.handle, pcm_full.data_ptr(), _lib.PCM_I16, S, 5120, pcm_full.stride(0),
                                    mel.data_ptr(), stream.cuda_stream))
    st = torch.zeros((2, S, 128), dtype=torch.float32, device=device)
    probs = torch.empty((S, FRAMES, 6), dtype=torch.float32, device=device)

    def gru_only(i):
        _lib.check(lib.kws_gru_forward(model.handle, mel.data_ptr(), S, FRAMES, None, st.data_ptr(), probs.data_ptr(),
                                       st.data_ptr(), None, stream.cuda_stream))

    def fe_only(i):
        _lib.check(lib.kws_frontend_mel(model.handle, pcm_full.data_ptr(), _lib.PCM_I16, S, 5120, pcm_full.stride(0),
                                        mel.data_ptr(), stream.cuda_stream))

    gru_only(0)
    k_iters = max(3, min(args.steps, 10))
    gru_ms, _ = timed_loop(torch, gru_only, k_iters, stream, lambda: None)
    fe_ms, _ = timed_loop(torch, fe_only, k_iters, stream, lambda: None)
    # roofline launch times: the kernels as they run inside the production step
    gru_launch_ms = parts[1] / cfg.num_layers
    fe_launch_ms = parts[0]
    peaks = load_peaks()
    flop_per_launch = S * FRAMES * GRU_FLOP_PER_FRAME / cfg.num_layers
    achieved_tf = flop_per_launch / (gru_launch_ms * 1e-3) / 1e12
    kname = "gru_tc_kernel" if args.precision == "tc" else "gru_layer_kernel"
    traffic = load_traffic(S) if args.precision == "tc" else {}
    gru_traffic = None
    if "gru_tc_kernel<0,1>" in traffic:
        gru_traffic = 0.5 * (traffic["gru_tc_kernel<0,1>"] + traffic["gru_tc_kernel<1,0>"])   # mean of the two launches
    roof_gru = dict(kernel=kname, bound="tensor", achieved=achieved_tf, peak=peaks["bf16_sustained"],
                    unit="TFLOP/s", frac=achieved_tf / peaks["bf16_sustained"], traffic=gru_traffic, launch_ms=gru_launch_ms,
                    peak_source=peaks["source"] + ", sustained bf16 (kernel timed inside a long step)",
                    note=("tcgen05 kind::f16 (fp16 operands, fp32 accumulate); " if args.precision == "tc" else
                          "exact fp32 FFMA kernel measured against the tensor-pipe peak; ") +
                         "algorithmic FLOP = 327,168 per frame (GRU 325,632 + FC 1,536), one launch per layer")
    fe_gbs = S * FE_BYTES_PER_STREAM_CHUNK / (fe_launch_ms * 1e-3) / 1e9
    roof_fe = dict(kernel="frontend_kernel", bound="hbm", achieved=fe_gbs, peak=peaks["hbm"], unit="GB/s",
                   frac=fe_gbs / peaks["hbm"], traffic=traffic.get("frontend_kernel"), launch_ms=fe_launch_ms,
                   algorithmic_bytes=S * FE_BYTES_PER_STREAM_CHUNK, traffic_source=traffic.get("_source"),
                   peak_source=peaks["source"] + ", copy bandwidth",
                   note="algorithmic bytes = %d per stream-chunk (int16 PCM in, carried tail r+w, fp32 mel out, flags); "
                        "one launch per step; timed inside the production step (fused VAD/tail work included)" % FE_BYTES_PER_STREAM_CHUNK)
    # the roofline of the dominant kernel (largest launch time); the other one alongside
    roofline, roofline_other = (roof_fe, roof_gru) if fe_launch_ms >= gru_launch_ms else (roof_gru, roof_fe)
    del pcm_full, mel, probs, st

    # ---- end to end through the host-buffer C-ABI call, served in waves
    e2e = None
    latency = None
    if not args.no_e2e:
        det.close()                                  # the full-batch server is not needed any more
        del chunks[2:]
        W = max(1, min(args.waves, S // 128))
        Sw = S // W
        assert Sw * W == S, "--streams must be divisible by --waves"
        n_host = 2
        host = [[torch.empty((Sw, CHUNK), dtype=torch.int16).pin_memory() for _ in range(n_host)] for _ in range(W)]
        for w in range(W):
            for b in range(n_host):
                host[w][b].copy_(chunks[b][w * Sw:(w + 1) * Sw])
        host_trig = [torch.zeros(Sw, dtype=torch.int32).pin_memory() for _ in range(W)]
        dets = [StreamingDetector(model, Sw) for _ in range(W)]
        wstreams = [torch.cuda.Stream(device=device) for _ in range(W)]
        lat = []

        def run_e2e(n_steps, record):
            """n_steps chunks for every stream: wave by wave, at most --depth waves in flight.  A wave's latency runs
            from the moment its chunk is handed to the C-ABI call (host clock) to its trigger flags being back on the host."""
            from collections import deque
            inflight = deque()
            for k in range(n_steps):
                for w in range(W):
                    if len(inflight) >= args.depth:
                        t_enq, ev = inflight.popleft()
                        ev.synchronize()
                        if record:
                            lat.append((time.perf_counter() - t_enq) * 1e3)
                    t_enq = time.perf_counter()
                    with torch.cuda.stream(wstreams[w]):
                        dets[w].step_host(host[w][k % n_host], host_trig[w])
                        ev = torch.cuda.Event()
                        ev.record(wstreams[w])
                    inflight.append((t_enq, ev))
            while inflight:
                t_enq, ev = inflight.popleft()
                ev.synchronize()
                if record:
                    lat.append((time.perf_counter() - t_enq) * 1e3)

        run_e2e(3, False)
        torch.cuda.synchronize()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record(stream)
        for ws in wstreams:
            ws.wait_stream(stream)
        run_e2e(args.steps, True)
        for ws in wstreams:
            stream.wait_stream(ws)
        ev1.record(stream)
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - t0) * 1e3
        barrier()
        e2e_ms = ev0.elapsed_time(ev1)
        e2e_ms_max = sharding.reduce_max_scalar(max(e2e_ms, 0.0), device)
        e2e = dict(value=world * S * AUDIO_S_PER_CHUNK * args.steps / (e2e_ms_max * 1e-3), unit=UNIT,
                   h2d_bytes_per_step=S * CHUNK * 2, d2h_bytes_per_step=S * 4, ms_per_step=e2e_ms_max / args.steps,
                   wall_ms_per_step=wall_ms / args.steps, waves=W, depth=args.depth,
                   note="pinned host int16 PCM -> H2D -> kws_stream_step -> D2H trigger flags for every stream every step, "
                        "through kws_stream_step_host; the streams are served as %d waves of %d (one stream object and CUDA "
                        "stream each), %d in flight, so H2D of one wave overlaps the kernels of another; bound by the host link "
                        "(%.1f GB/s achieved)" % (W, Sw, args.depth, S * CHUNK * 2 / (e2e_ms_max / args.steps * 1e-3) / 1e9))
        lat.sort()
        latency = dict(unit="ms", p50=lat[len(lat) // 2], p99=lat[min(len(lat) - 1, int(0.99 * len(lat)))], max=lat[-1],
                       samples=len(lat),
                       definition="host clock from handing a wave's 300 ms chunk (pinned host memory) to kws_stream_step_host "
                                  "until its trigger flags are back on the host, measured inside the e2e timed region "
                                  "(includes queueing behind the %d wave(s) in flight)" % (args.depth - 1))
        trig_total += int(sum(int(t.sum()) for t in host_trig))
        for d in dets:
            d.close()
        del host
    clocks = sampler.stop() if rank == 0 else None

    if world > 1:
        dist.all_reduce(trig_total)                  # the final result gather (NCCL), off the hot path
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_port_throughput(*cpu_sample_size(args, budget_s=15.0))
    if rank == 0:
        per_sorted = sorted(per_step)
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
                    ms_per_step=ms_per_step, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype="f16 operands / f32 accumulate+state" if args.precision == "tc" else "f32",
                    data="synthetic",
                    config=dict(workload="streaming serve (BASELINE configs[2]): %d concurrent streams per GPU with carried GRU "
                                         "state and VAD reset, 300 ms chunks, 2L GRU-128 n_mel=40 6 classes" % S,
                                streams_per_gpu=S, chunk_samples=CHUNK, frames_per_chunk=FRAMES, sharding="streams/dp%d" % world,
                                l2_policy="inputs larger than L2: %d rotating 1.26 GB PCM buffers" % n_buf),
                    realtime_streams=value / 1.0,
                    realtime_streams_e2e=(e2e["value"] if e2e else None),
                    chunk_latency_ms_p99=(latency["p99"] if latency else None), chunk_latency=latency,
                    full_batch_step_ms_p99=per_sorted[min(len(per_sorted) - 1, int(0.99 * len(per_sorted)))],
                    kernels_ms=dict(frontend=parts[0], gru_2_layers=parts[1], decode=parts[2], step_total=ms_per_step,
                                    standalone_frontend_no_pre_step=fe_ms / k_iters, standalone_gru_row_major_mel=gru_ms / k_iters,
                                    note="frontend/gru_2_layers/decode: CUDA events inside kws_stream_step (debug hook), "
                                         "mean of 6 synchronised steps after the timed region"),
                    roofline=roofline, roofline_other=[roofline_other], cpu_baseline=cpu, e2e=e2e,
                    gpu_launches=4 * args.steps, clocks=clocks, host_binding=host_binding,
                    triggers=int(trig_total.item()))
        print(json.dumps(line), flush=True)
    det.close()
    model.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
