#!/usr/bin/env python3
"""bench.py -- throughput of the in_cwave signal chain (Hilbert -> modulator graph -> render) on B200.

    python bench.py --gpus N --steps K --warmup W [--workload c4|c2|c1|c3] [--impl reference]

One "step" = one pass of the hot path over one batch of synthetic input of the named workload.
Prints ONE JSON line (rank 0).  Multi-GPU: one process per GPU under torch.distributed.run; the
path shards by independent streams, no data-path collective, "scaling": "weak" (every rank runs
the whole per-GPU workload; value = frames of all ranks / max-over-ranks device time).

  value      stereo Mframes/s with inputs resident in HBM (CUDA events around the timed steps)
  e2e        the same through the C ABI's host entry point: pinned host buffers, H2D + kernels + D2H
  roofline   HBM roofline of the dominant kernel (per-kernel CUDA events on the launching stream):
             algorithmic bytes of the path per launch / that kernel's average launch duration
  cpu_baseline  the reference's own C code (oracle/_ref) or our port of it, timed on host cores

`--impl reference` times the reference's CPU implementation alone, on all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from in_cwave_b200 import spec as S  # noqa: E402
from in_cwave_b200 import synth  # noqa: E402

METRIC = "stereo Mframes/s, Hilbert + modulator graph + 24-bit render (2x for channel-samples/s)"
UNIT = "Mframes/s"


# ---------------------------------------------------------------------------------------------
# workloads (BASELINE.json configs)
# ---------------------------------------------------------------------------------------------
def workload(name: str) -> dict:
    if name == "c4":
        # 4096 independent 48 kHz stereo f32 streams, 10 s each; C1 graph; reference-exact Hilbert
        return dict(name="c4", desc="4096 x 48 kHz stereo f32 WAV, 10 s each: Hilbert(T1,Kahan,reject) + 100 Hz shift + 24-bit render, no dither",
                    spec=S.config_c1(hilbert_mode="exact"), streams=4096, frames=480_000, chunk=480_000,
                    bytes_per_frame=14, hilbert="exact")
    if name == "c4w":
        # the same batch on EVERY rank (N x 4096 streams in all): the exact kernel's run time is set by the 480 000 serial samples of a
        # stream, not by the stream count, so c4 cut by stream does not speed up with N -- this line shows what N GPUs carry in that time
        w = workload("c4")
        w.update(name="c4w", weak_batch=True,
                 desc="4096 x 48 kHz stereo f32 WAV per GPU, 10 s each (the c4 batch on every rank): Hilbert(T1,Kahan,reject) + 100 Hz shift + 24-bit render, no dither")
        return w
    if name == "c4ns":
        # C4 with a noise shaper: the quantiser's error feedback is serial per channel
        return dict(name="c4ns", desc="4096 x 44.1 kHz stereo f32 WAV, 10 s each: Hilbert(T1,Kahan,reject) + 100 Hz shift + TPDF + "
                                      "20-tap Shibata noise shaper, 24-bit render",
                    spec=S.config_c1(hilbert_mode="exact", sample_rate=44100, render_type=2, nshape_type=6), streams=4096,
                    frames=441_000, chunk=16_384, bytes_per_frame=14, hilbert="exact")
    if name == "c1":
        return dict(name="c1", desc="48 kHz stereo f32 WAV, 60 s: Hilbert + 100 Hz shift + 24-bit render (one stream)",
                    spec=S.config_c1(hilbert_mode="scan"), streams=1, frames=2_880_000, chunk=2_880_000,
                    bytes_per_frame=14, hilbert="scan")
    if name == "c3":
        return dict(name="c3", desc="CWAVE f32 I/Q 96 kHz stereo, 600 s: 2 shifts + PM + mix, 16-bit render",
                    spec=S.config_c3(), streams=1, frames=57_600_000, chunk=57_600_000,
                    bytes_per_frame=20, hilbert="none")
    if name == "c2":
        return dict(name="c2", desc="192 kHz 24-bit PCM stereo, 1 h: Hilbert + 100 Hz shift + TPDF dither + 24-bit render (one stream)",
                    spec=S.config_c2(hilbert_mode="scan"), streams=1, frames=691_200_000, chunk=691_200_000,
                    bytes_per_frame=12, hilbert="scan")
    if name == "c5":
        # one 384 kHz stereo f32 stream cut in time across the ranks: 3 h (1/8 of 24 h) per GPU,
        # NCCL hand-off of the Hilbert state at every cut (in_cwave_b200.dist.run_time_sharded)
        return dict(name="c5", desc="384 kHz stereo f32, 24 h stream time-sharded: 3 h per GPU, Hilbert + 100 Hz shift + 24-bit render, "
                                    "NCCL filter-state hand-off between neighbours",
                    spec=S.config_c1(hilbert_mode="scan", sample_rate=384000), streams=1, frames=4_147_200_000,
                    chunk=4_147_200_000, bytes_per_frame=14, hilbert="scan", time_sharded=True)
    raise SystemExit(f"unknown workload {name}")


def peaks() -> dict:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm_gbs=float(d["hbm_gbs"]), source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, source="fallback (B200_PROFILING.md)")


# ---------------------------------------------------------------------------------------------
# clocks during the timed region
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=3)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


# ---------------------------------------------------------------------------------------------
# CPU legs: the reference's own code on host cores
# ---------------------------------------------------------------------------------------------
def _cpu_worker(args):
    """One process = one reference plugin instance (process-global state): run its share of streams.
    Only the reference/oracle call is on the clock; input synthesis is not."""
    spec, n_frames, stream_ids, use_ref = args
    from oracle import pyoracle as po
    frames, busy = 0, 0.0
    for sid in stream_ids:
        raw = synth.stream_bytes(spec, n_frames, stream_id=sid)
        t1 = time.perf_counter()
        if use_ref:
            out = po.ref_process(spec, raw, read_quant=4096, tmpdir=os.environ.get("ICW_TMPDIR"))
        else:
            out = po.port_process(spec, raw)
        busy += time.perf_counter() - t1
        frames += out["frames"]
    return frames, busy


def cpu_run(spec: dict, n_frames: int, streams_per_core: int, cores: int):
    """cores processes, each `streams_per_core` fresh streams of n_frames; returns (frames, seconds, kind).
    Input synthesis is inside the workers but outside the reference's clock below (measured apart)."""
    import multiprocessing as mp
    from oracle import pyoracle as po
    use_ref = po.have_ref()
    if not use_ref:
        po.port()
    jobs = [(spec, n_frames, [c * streams_per_core + i for i in range(streams_per_core)], use_ref) for c in range(cores)]
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    frames = sum(r[0] for r in res)
    busy = max(r[1] for r in res)
    return frames, max(busy, 1e-9), wall, ("reference" if use_ref else "port")


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def bind_near_gpu(index: int) -> str:
    """Multi-GPU runs: keep this rank's threads (and so its pinned staging pages, first touch) on the
    CPUs next to its GPU.  Returns a note for the JSON line; never fatal."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} CPUs local to GPU {index}"
    except Exception as ex:
        return f"not bound ({type(ex).__name__})"
    return "not bound"


def cpu_model() -> str:
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def base_config(wl: dict) -> dict:
    """The keys both arms print under `config`: what the workload IS (the arms differ in how they run it)."""
    return dict(workload=wl["desc"], streams=wl["streams"], frames_per_stream=wl["frames"])


def reference_arm(args, wl):
    """--impl reference: the reference's CPU implementation of the path, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    total_steps = args.steps + args.warmup
    budget = max(2.0, 150.0 / max(1, total_steps))           # seconds per step
    spec = wl["spec"]
    frames = min(wl["frames"], 480_000)
    per_core = 1.0e6 * budget                                   # ~1.0 Mframes/s/core for this chain
    spc = int(per_core // frames)
    if spc < 1:
        spc, frames = 1, max(20_000, int(per_core))
    times, fr = [], 0
    kind = "port"
    for i in range(total_steps):
        f, busy, wall, kind = cpu_run(spec, frames, spc, cores)
        if i >= args.warmup:
            times.append(busy)
            fr = f
    ms = 1e3 * float(np.mean(times))
    val = fr / (ms * 1e-3) / 1e6
    sample = f"{cores} processes x {spc} fresh streams x {frames} frames per step ({wl['name']} chain, same spec)"
    line = dict(impl="reference", metric=METRIC, value=val, unit=UNIT, n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f64", data="synthetic",
                config=base_config(wl),
                parity=dict(hilbert="reference: serial DF-II recurrences on the CPU (the definition of bit-exact)"),
                cpu_baseline=dict(value=val, unit=UNIT, cores=cores, kind=kind, sample=sample, cpu=cpu_model()),
                e2e=dict(value=val, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


class _NoDist:
    """world == 1: the time-sharded path has nobody to talk to."""
    class ReduceOp:
        SUM = MAX = None

    @staticmethod
    def batch_isend_irecv(ops):
        return []

    @staticmethod
    def all_reduce(t, op=None, group=None):
        return None


# ---------------------------------------------------------------------------------------------
# in-run parity: the bytes this run produced against the reference's own code on the same input
# ---------------------------------------------------------------------------------------------
def _oracle_pcm(spec: dict, raw: np.ndarray):
    """PCM of a FRESH reference context over raw (oracle/_ref when it travelled with the repo, else the port)."""
    from oracle import pyoracle as po
    if po.have_ref():
        r = po.ref_process(spec, raw, read_quant=4096, tmpdir=os.environ.get("ICW_TMPDIR"))
        _oracle_pcm.last_rejects = int(r["stats"].subnorm_cnt)      # the reference's "subnorm reject" hits (src/hblpf.c:915,1046)
        return r["pcm"], "reference"
    _oracle_pcm.last_rejects = None
    return po.port_process(spec, raw)["pcm"], "port"


def _pcm_ints(pcm: np.ndarray, bps: int) -> np.ndarray:
    b = np.asarray(pcm, dtype=np.uint8).reshape(-1, bps).astype(np.int64)
    v = b[:, 0] | (b[:, 1] << 8)
    if bps == 3:
        v |= b[:, 2] << 16
        return np.where(v >= 1 << 23, v - (1 << 24), v)
    return np.where(v >= 1 << 15, v - (1 << 16), v)


def pcm_distance(got: np.ndarray, want: np.ndarray, bps: int) -> dict:
    g, w = _pcm_ints(got, bps), _pcm_ints(want, bps)
    d = np.abs(g - w)
    rms_w = float(np.sqrt(np.mean(w.astype(np.float64) ** 2))) or 1.0
    return dict(samples=int(g.size), mismatches=int(np.count_nonzero(d)), max_lsb=int(d.max() if d.size else 0),
                rms_vs_reference=float(np.sqrt(np.mean(d.astype(np.float64) ** 2))) / rms_w)


def parity_check(wl: dict, d_in, d_out, K: int, N: int) -> dict:
    """Sampled windows of the LAST timed step's output (every step starts from fresh contexts) against the reference:
    one long stream -> its first 2^18 frames; a batch -> three whole streams."""
    spec = wl["spec"]
    fb, ob = S.frame_bytes(spec), S.out_frame_bytes(spec)
    bps = ob // 2
    # the reference's Hilbert is always its own serial recurrence: compare against the exact-mode spec
    ref_spec = dict(spec, hilbert_mode="exact")
    t0 = time.perf_counter()
    if K == 1:
        W = min(N, 1 << 18)
        picks = [(0, W)]
    else:
        W = min(N, 480_000)
        picks = [(k, W) for k in sorted({0, K // 3, K - 1})]
    tot = dict(samples=0, mismatches=0, max_lsb=0)
    sq, kind = 0.0, "port"
    ref_rejects = 0
    for k, w in picks:
        raw = d_in[k, : w * fb].cpu().numpy()
        got = d_out[k, : w * ob].cpu().numpy()
        want, kind = _oracle_pcm(ref_spec, raw)
        ref_rejects = None if (_oracle_pcm.last_rejects is None or ref_rejects is None) else ref_rejects + _oracle_pcm.last_rejects
        r = pcm_distance(got, want, bps)
        tot["samples"] += r["samples"]; tot["mismatches"] += r["mismatches"]; tot["max_lsb"] = max(tot["max_lsb"], r["max_lsb"])
        sq += r["rms_vs_reference"] ** 2 * r["samples"]
    tot["rms_vs_reference"] = float(np.sqrt(sq / max(1, tot["samples"])))
    tot["against"] = f"oracle/{'_ref (the compiled reference)' if kind == 'reference' else 'port'}, fresh context per stream"
    tot["window"] = (f"first {picks[0][1]} frames of the stream" if K == 1 else
                     f"streams {[k for k, _ in picks]} x {W} frames (whole streams)")
    tot["hilbert"] = wl["hilbert"]
    tot["seconds"] = round(time.perf_counter() - t0, 2)
    if wl["hilbert"] == "scan" and int(spec.get("is_subnorm_reject", 1)) and not str(spec.get("fmt", "")).startswith("cw_"):
        # the reference zeroes a filter state below 1.0 (its flag compared as a number): that only fires on a state that is
        # (almost) empty -- the first samples of a stream, digital silence.  The modal scan has no such test: counted here.
        tot["subnorm_rejects"] = dict(reference=ref_rejects, ours=0,
                                      note="scan mode does not model the reference's |w| < 1 state zeroing (start-up / digital "
                                           "silence only); exact mode counts the same hits as the reference (tests/test_gpu_parity.py)")
    if wl["hilbert"] == "scan":
        tot["note"] = ("scan mode evaluates the filter exactly (3e-15 of binary128); the reference's serial FP64 recurrence carries "
                       "its own rounding noise (about 5e-5 of RMS for the default design), which is the distance counted here")
    return tot


# ---------------------------------------------------------------------------------------------
# one workload on this rank: resident timing, per-kernel roofline, e2e, parity
# ---------------------------------------------------------------------------------------------
# FP64 instructions per stereo frame the kernels that run these workloads actually issue: the one-kernel
# scan path 2 x 73 DFMA per frame (c2: icw_sfused.cu) or the three-pass scan 164 + 82 (c1, and c5: no dither), the exact Kahan recurrences
# 4 x 281 (c4), plus the frame path (oscillator, DSP list, dither, quantiser)
FP64_OPS = {"c2": 146 + 105, "c5": 164 + 82 + 70, "c1": 164 + 82 + 70, "c4": 1124 + 70, "c4w": 1124 + 70, "c4ns": 1124 + 200, "c3": 300}


def measure(ctx, name: str, steps: int, warmup: int, want_e2e: bool, want_parity: bool, args=None) -> dict:
    import torch
    from in_cwave_b200 import dist as D
    dist, world, rank, local, dev, eng = ctx["dist"], ctx["world"], ctx["rank"], ctx["local"], ctx["dev"], ctx["eng"]
    wl = workload(name)
    if args is not None and args.streams:
        wl["streams"] = args.streams
    if args is not None and args.frames:
        wl["frames"] = args.frames
        wl["chunk"] = min(wl["chunk"], args.frames)
    spec, N, chunk = wl["spec"], wl["frames"], wl["chunk"]
    K_all = wl["streams"]
    # how the workload spreads over the ranks: a batch is cut by stream (fixed total: strong scaling);
    # one long stream is cut in time (c5) or, having nothing to cut, replicated (weak scaling, no collective)
    if K_all > 1 and world > 1 and not wl.get("weak_batch"):
        lo, hi = D.shard_streams(K_all, rank, world)
        K, scaling, sharding = hi - lo, "strong", f"by stream: {K_all} streams / {world} ranks, no collective"
        frames_step = K_all * N
    elif wl.get("time_sharded"):
        K, scaling = 1, "weak"
        sharding = f"by time: rank r plays frames [r*{N}, (r+1)*{N}) of ONE stream of {world * N} frames; NCCL hand-off of the Hilbert state"
        frames_step = N * world
    else:
        K, scaling = K_all, "weak"
        sharding = ("single GPU" if world == 1 else f"by stream: {K_all} streams on every rank, {K_all * world} in all, no collective" if K_all > 1
                    else "replicas: one independent stream per rank, no collective")
        frames_step = K * N * world
    fb, ob = S.frame_bytes(spec), S.out_frame_bytes(spec)
    ses = eng.session(spec, K)
    d_in = synth.device_fill(spec, K, N, dev)
    in_stride = d_in.stride(0)
    out_stride = (N * ob + 15) // 16 * 16
    d_out = torch.empty((K, out_stride), dtype=torch.uint8, device=dev)
    cs = torch.cuda.current_stream().cuda_stream
    shard_be = None
    if wl.get("time_sharded"):
        if world > 1 and ctx.get("comm") is None:
            ctx["comm"] = D.make_comm(dist, rank, world, dev)     # our own ncclComm_t; the id travels over torch.distributed
        shard_be = D.CudaBackend(eng, spec, comm=ctx.get("comm"))
        shard_be.ses.close()
        shard_be.ses = ses                            # profile / count launches on the session bench reads
    handoff_ms = []

    def one_step():
        ses.reset()                                   # every step = the same fresh streams
        if shard_be is not None:
            info = D.run_time_sharded(shard_be, dist if world > 1 else _NoDist(), spec, d_in[0, : N * fb], rank * N, rank, world,
                                      device=dev, d_out=d_out[0])
            if isinstance(info, tuple) and len(info) > 3 and info[3] is not None:
                handoff_ms.append(info[3])
            return
        for f0 in range(0, N, chunk):
            n = min(chunk, N - f0)
            ses.process_device(d_in.data_ptr() + f0 * fb, n, d_out.data_ptr() + f0 * ob,
                               in_stride=in_stride, out_stride=out_stride, stream=cs)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        one_step()
    barrier()
    handoff_ms.clear()
    launches0 = ses.stats()["kernel_launches"]
    ses.profile(True)
    ses.profile_read(reset=True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(steps):
        one_step()
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    prof = ses.profile_read(reset=True)
    ses.profile(False)
    launches = ses.stats()["kernel_launches"] - launches0
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / steps
    value = frames_step / (ms_step * 1e-3) / 1e6

    # ---- roofline of the dominant kernel (this rank's launches) -----------------------------------------
    pk = peaks()
    dom = max(prof, key=lambda k: prof[k]["ms"])
    spans = max(1, prof[dom]["launches"])
    dom_ms = prof[dom]["ms"] / spans                         # one span = the class's kernels over one launch group
    units_per_launch = K * N * steps / spans                 # frames one such span processes
    algo_bytes = units_per_launch * wl["bytes_per_frame"]
    achieved = algo_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    traffic = None
    for tf in ("r2_traffic.json", "r1_traffic.json"):        # dram bytes per frame from the ncu --set full captures
        tfile = ROOT / "profiles" / tf
        if tfile.exists():
            tt = json.loads(tfile.read_text()).get("c4" if wl["name"] == "c4w" else wl["name"], {}).get(dom)
            if tt:
                traffic = tt["dram_bytes_per_frame"] * units_per_launch
                break
    fp64_ops = FP64_OPS.get(wl["name"], 0)
    total_ms = max(1e-9, sum(x["ms"] for x in prof.values()))
    roofline = dict(bound="hbm", kernel=dom, achieved=achieved, peak=pk["hbm_gbs"], unit="GB/s",
                    frac=achieved / pk["hbm_gbs"], traffic=traffic, peak_source=pk["source"],
                    algorithmic_bytes_per_frame=wl["bytes_per_frame"], frames_per_launch=units_per_launch,
                    avg_launch_ms=dom_ms,
                    whole_step_frac=(K * N * wl["bytes_per_frame"] / (ms_step * 1e-3) / 1e9) / pk["hbm_gbs"],
                    kernel_share={k: v["ms"] / total_ms for k, v in prof.items()},
                    binding_bound=dict(kind="fp64 pipe / instruction issue (not HBM)", peak_tops=18.55, peak_tops_3_vector_operands=12.37,
                                       peak_source="tools/fp64_probe.cu on this pool's B200 (profiles/r1_fp64_probe_b200.txt)",
                                       approx_ops_per_frame=fp64_ops,
                                       whole_step_frac=(K * N) * fp64_ops / (ms_step * 1e-3) / 18.55e12))

    res = dict(desc=wl["desc"], value=value, unit=UNIT, ms_per_step=ms_step, steps=steps, warmup=warmup, scaling=scaling,
               sharding=sharding, hilbert=wl["hilbert"], streams_this_rank=K, frames_per_stream=N, frames_per_step=frames_step,
               roofline=roofline, gpu_launches=int(launches), kernel_ms={k: v["ms"] / steps for k, v in prof.items()},
               clocks=clocks, l2="inputs larger than L2 (per-step input %.2f GB on this rank)" % (K * N * fb / 1e9),
               config=base_config(wl))
    if handoff_ms:
        hm = torch.tensor([float(np.mean(handoff_ms))], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(hm, op=dist.ReduceOp.MAX)
        res["handoff_ms"] = float(hm.item())
        res["handoff"] = ("icw_session_handoff (C ABI): ncclSend/ncclRecv of the %d-double filter state, device to device, inside the "
                          "timed region; NCCL %s" % (4 * 20, _abi_nccl_version()))

    # ---- parity of what the timed steps wrote ---------------------------------------------------------
    if want_parity:
        pc = None
        if rank == 0:
            try:
                pc = parity_check(wl, d_in, d_out, K, N)
            except Exception as ex:                           # the bench line must still appear
                pc = dict(error=f"{type(ex).__name__}: {ex}"[:300])
        if shard_be is not None and world > 1:
            st = stitch_check(ctx, wl, d_in, d_out, N)
            if rank == 0:
                pc = dict(pc or {}, stitch=st)
        res["parity_check"] = pc

    # ---- end to end through the host entry point --------------------------------------------------------
    if want_e2e and shard_be is None:
        try:
            h_in = torch.empty((K, N * fb), dtype=torch.uint8, pin_memory=True)
            h_out = torch.empty((K, N * ob), dtype=torch.uint8, pin_memory=True)
            h_in.copy_(d_in[:, : N * fb])
            torch.cuda.synchronize()

            def e2e_step():
                ses.reset()
                for f0 in range(0, N, chunk):
                    n = min(chunk, N - f0)
                    ses.process_host_into(h_in.data_ptr() + f0 * fb, N * fb, n, h_out.data_ptr() + f0 * ob, N * ob)

            e2e_step()
            barrier()
            t0 = time.perf_counter()
            reps = max(1, min(steps, 2))
            for _ in range(reps):
                e2e_step()
            barrier()
            dt = (time.perf_counter() - t0) / reps
            tt = torch.tensor([dt], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            same = bool(torch.equal(h_out[0, : min(N, 1 << 16) * ob].to(dev), d_out[0, : min(N, 1 << 16) * ob]))
            # the ceiling of that number on this host: the same bytes over the same pinned buffers with NO kernels, H2D and
            # D2H at once on two streams, every rank at the same time (name the end-to-end limiter)
            s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
            copy_s = []
            for _ in range(2):
                barrier()
                t0 = time.perf_counter()
                with torch.cuda.stream(s_in):
                    d_in[:, : N * fb].copy_(h_in, non_blocking=True)
                with torch.cuda.stream(s_out):
                    h_out.copy_(d_out[:, : N * ob], non_blocking=True)
                s_in.synchronize()
                s_out.synchronize()
                barrier()
                copy_s.append(time.perf_counter() - t0)
            tc = torch.tensor([min(copy_s)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tc, op=dist.ReduceOp.MAX)
            copy_ms = float(tc.item()) * 1e3
            res["e2e"] = dict(value=frames_step / float(tt.item()) / 1e6, unit=UNIT, h2d_bytes_per_step=K * N * fb,
                              d2h_bytes_per_step=K * N * ob, ms_per_step=float(tt.item()) * 1e3, steps=reps,
                              how="icw_session_process_host on pinned host buffers; H2D + kernels + D2H inside the timed region",
                              host_affinity=ctx["numa_note"], same_bytes_as_resident_run=same,
                              copy_only=dict(ms=copy_ms, h2d_gbs_per_rank=K * N * fb / copy_ms / 1e6, d2h_gbs_per_rank=K * N * ob / copy_ms / 1e6,
                                             aggregate_gbs_each_way=world * K * N * max(fb, ob) / copy_ms / 1e6,
                                             e2e_over_copy_only=float(tt.item()) * 1e3 / copy_ms,
                                             how="the step's bytes, pinned <-> device, both directions at once, no kernels, all ranks "
                                                 "together: what the host's PCIe / memory path gives this many ranks"))
            del h_in, h_out
        except RuntimeError as ex:
            res["e2e"] = dict(value=None, unit=UNIT, error=str(ex)[:200])
    res["_spec"] = spec
    ses.close()
    del d_in, d_out
    torch.cuda.empty_cache()
    return res


def _abi_nccl_version() -> str:
    from in_cwave_b200 import _abi
    v = _abi.lib().icw_nccl_version()
    return f"{v // 10000}.{v // 100 % 100}.{v % 100}" if v else "unavailable"


def stitch_check(ctx, wl: dict, d_in, d_out, N: int) -> dict:
    """c5: is the cut invisible?  Every rank r > 0 replays, alone, the 2^19 frames before its cut from a zero filter
    state (the input is periodic over the ranks, so those are the last frames of its own buffer) and then the first
    2^16 frames of its shard, and compares them with what the time-sharded run wrote there after the NCCL hand-off."""
    import torch
    from in_cwave_b200 import dist as D
    dist, world, rank, dev, eng = ctx["dist"], ctx["world"], ctx["rank"], ctx["dev"], ctx["eng"]
    spec = wl["spec"]
    fb, ob = S.frame_bytes(spec), S.out_frame_bytes(spec)
    W, C = min(N, D.WARMUP_FRAMES), min(N, 1 << 16)
    stats = torch.zeros(3, dtype=torch.float64, device=dev)     # mismatching samples, max LSB, samples
    if rank > 0:
        be = D.CudaBackend(eng, spec)
        a = rank * N
        be.start_at(D.closed_form_state(spec, a - W), np.zeros(D.STATE_DOUBLES))
        scratch = torch.empty(W * ob + 16, dtype=torch.uint8, device=dev)
        be.process(d_in[0, (N - W) * fb: N * fb], scratch)
        again = torch.empty(C * ob + 16, dtype=torch.uint8, device=dev)
        be.process(d_in[0, : C * fb], again)
        torch.cuda.synchronize()
        r = pcm_distance(again[: C * ob].cpu().numpy(), d_out[0, : C * ob].cpu().numpy(), ob // 2)
        stats = torch.tensor([r["mismatches"], r["max_lsb"], r["samples"]], dtype=torch.float64, device=dev)
        be.ses.close()
    mx = stats.clone()
    dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    return dict(cuts=world - 1, samples=int(stats[2].item()), mismatches=int(stats[0].item()), max_lsb=int(mx[1].item()),
                how=f"each rank > 0: single-GPU replay of {W} warm-up frames + the first {C} frames of its shard vs the sharded run's bytes")


# ---------------------------------------------------------------------------------------------
# the GPU arm
# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=os.environ.get("ICW_WORKLOAD", "c2"))
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--streams", type=int, default=0, help="override the workload's stream count")
    ap.add_argument("--frames", type=int, default=0, help="override frames per stream")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-workloads", action="store_true", help="only the headline workload, no `workloads` object")
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()
    wl = workload(args.workload)
    if args.warmup < 3:
        args.warmup = 3

    if args.impl == "reference":
        if args.streams:
            wl["streams"] = args.streams
        if args.frames:
            wl["frames"] = args.frames
        reference_arm(args, wl)
        return

    import torch
    import torch.distributed as dist
    import in_cwave_b200 as icw

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    numa_note = bind_near_gpu(local) if world > 1 else "single rank: not bound"
    eng = icw.Engine(local)
    ctx = dict(dist=dist, world=world, rank=rank, local=local, dev=dev, eng=eng, numa_note=numa_note)

    # ---- the headline workload: full treatment --------------------------------------------------------
    head = measure(ctx, args.workload, args.steps, args.warmup, want_e2e=not args.no_e2e, want_parity=not args.no_parity, args=args)
    spec = head.pop("_spec")

    # ---- the other named configs, same run, same box: the bit-exact batch sharded by stream, the full
    # graph, and -- where there is more than one GPU -- the time-sharded stream with its NCCL hand-off timed ------------
    others = {}
    if not args.no_workloads and args.workload == "c2" and not args.streams and not args.frames:
        for name in ["c4", "c3", "c4ns"] + (["c5", "c4w"] if world > 1 else []):
            try:
                r = measure(ctx, name, max(3, min(args.steps, 10)), 3, want_e2e=not args.no_e2e and name != "c4w", want_parity=not args.no_parity)
                r.pop("_spec", None)
                others[name] = r
            except Exception as ex:
                others[name] = dict(error=f"{type(ex).__name__}: {ex}"[:300])

    # ---- CPU baseline on this box's host cores (rank 0, N = 1 only) ------------------------------------
    cpu = None
    if wl.get("time_sharded"):
        args.no_cpu = True
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = host_cores()
        fr = min(head["frames_per_stream"], 480_000)
        spc = max(1, int(12.0e6 // fr))               # ~12 s of single-core work per process
        f, busy, wall, kind = cpu_run(spec, fr, spc, cores)
        cpu = dict(value=f / busy / 1e6, unit=UNIT, cores=cores, kind=kind, cpu=cpu_model(),
                   sample=f"{cores} processes x {spc} fresh streams x {fr} frames of the same chain ({busy:.1f} s busy)")

    if ctx.get("comm") is not None:
        from in_cwave_b200 import dist as D
        D.free_comm(ctx["comm"])
    if rank == 0:
        hil = head["hilbert"]
        line = dict(metric=METRIC, value=head["value"], unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=head["ms_per_step"], higher_is_better=True, scaling=head["scaling"], vs_baseline=None, dtype="f64",
                    data="synthetic",
                    parity=dict(hilbert=hil,
                                note=("exact: bit-exact with the reference (tests/test_gpu_parity.py)" if hil != "scan" else
                                      "scan: analytic signal within 1e-12 of the binary128 evaluation of the reference's filter; "
                                      "everything downstream byte-exact; the reference's own FP64 rounding noise (5e-5 of RMS for the "
                                      "default design) separates its PCM from ours -- counted in parity_check; the bit-exact mode is "
                                      "workloads.c4 (tests/test_gpu_scan.py)")),
                    parity_check=head.get("parity_check"),
                    config=head["config"],
                    timing=dict(streams_per_gpu=head["streams_this_rank"], frames_per_launch=head["roofline"]["frames_per_launch"],
                                sharding=head["sharding"], l2=head["l2"]),
                    roofline=head["roofline"], cpu_baseline=cpu, e2e=head.get("e2e"), gpu_launches=head["gpu_launches"],
                    clocks=head["clocks"], kernel_ms=head["kernel_ms"])
        if "handoff_ms" in head:
            line["handoff_ms"] = head["handoff_ms"]
        if others:
            line["workloads"] = others
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
