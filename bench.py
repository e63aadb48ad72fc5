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
                    spec=S.config_c1(hilbert_mode="exact"), streams=4096, frames=480_000, chunk=16_384,
                    bytes_per_frame=14, hilbert="exact")
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
                config=dict(workload=wl["desc"], hilbert=wl["hilbert"]),
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
    args = ap.parse_args()
    wl = workload(args.workload)
    if args.streams:
        wl["streams"] = args.streams
    if args.frames:
        wl["frames"] = args.frames
        wl["chunk"] = min(wl["chunk"], args.frames)
    if args.warmup < 3:
        args.warmup = 3

    if args.impl == "reference":
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

    spec, K, N, chunk = wl["spec"], wl["streams"], wl["frames"], wl["chunk"]
    fb, ob = S.frame_bytes(spec), S.out_frame_bytes(spec)
    eng = icw.Engine(local)
    ses = eng.session(spec, K)

    # resident synthetic input, one row per stream (rows padded to 16 B)
    d_in = synth.device_fill(spec, K, N, dev)
    in_stride = d_in.stride(0)
    out_stride = (N * ob + 15) // 16 * 16
    d_out = torch.empty((K, out_stride), dtype=torch.uint8, device=dev)
    cs = torch.cuda.current_stream().cuda_stream

    shard_be = None
    if wl.get("time_sharded"):
        from in_cwave_b200 import dist as D
        shard_be = D.CudaBackend(eng, spec)
        shard_be.ses.close()
        shard_be.ses = ses                            # profile / count launches on the session bench reads

    def one_step():
        ses.reset()                                   # every step = the same fresh streams
        if shard_be is not None:
            # rank r plays frames [r*N, (r+1)*N) of one stream of world*N frames
            D.run_time_sharded(shard_be, dist if world > 1 else _NoDist(), spec, d_in[0, : N * fb], rank * N, rank, world,
                               device=dev, d_out=d_out[0])
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

    for _ in range(args.warmup):
        one_step()
    barrier()
    launches0 = ses.stats()["kernel_launches"]
    ses.profile(True)
    ses.profile_read(reset=True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
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
    ms_step = float(t.item()) / args.steps
    frames_step = K * N * world
    value = frames_step / (ms_step * 1e-3) / 1e6

    # ---- roofline of the dominant kernel --------------------------------------------------------
    pk = peaks()
    dom = max(prof, key=lambda k: prof[k]["ms"])
    spans = max(1, prof[dom]["launches"])
    dom_ms = prof[dom]["ms"] / spans                         # one span = the class's kernels over one launch group
    units_per_launch = K * N * args.steps / spans            # frames one such span processes
    algo_bytes = units_per_launch * wl["bytes_per_frame"]
    achieved = algo_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    traffic = None
    tfile = ROOT / "profiles" / "r1_traffic.json"            # dram bytes per frame from the ncu --set full captures
    if tfile.exists():
        t = json.loads(tfile.read_text()).get(wl["name"], {}).get(dom)
        if t:
            traffic = t["dram_bytes_per_frame"] * units_per_launch
    # FP64 thread-operations per frame (counted from the SASS of the sample loops) and the
    # measured issue rate of the FP64 pipe (tools/fp64_probe.cu): 64 lanes/clk/SM when an instruction reads two
    # vector registers (the third operand uniform), 42.7 when it reads three -- most of the modal DFMAs do
    fp64_ops = {"c2": 164 + 82 + 105, "c5": 164 + 82 + 70, "c1": 164 + 82 + 70, "c4": 1124 + 70, "c4ns": 1124 + 200, "c3": 300}.get(wl["name"], 0)
    roofline = dict(bound="hbm", kernel=dom, achieved=achieved, peak=pk["hbm_gbs"], unit="GB/s",
                    frac=achieved / pk["hbm_gbs"], traffic=traffic, peak_source=pk["source"],
                    algorithmic_bytes_per_frame=wl["bytes_per_frame"], frames_per_launch=units_per_launch,
                    avg_launch_ms=dom_ms,
                    kernel_share={k: v["ms"] / max(1e-9, sum(x["ms"] for x in prof.values())) for k, v in prof.items()},
                    binding_bound=dict(kind="fp64 pipe / instruction issue (not HBM)", peak_tops=18.55, peak_tops_3_vector_operands=12.37,
                                       peak_source="tools/fp64_probe.cu on this pool's B200 (profiles/r1_fp64_probe_b200.txt)",
                                       approx_ops_per_frame=fp64_ops,
                                       whole_step_frac=(frames_step / world) * fp64_ops / (ms_step * 1e-3) / 18.55e12))

    # ---- end to end through the host entry point ----------------------------------------------------
    e2e = None
    if not args.no_e2e and not wl.get("time_sharded"):
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
            reps = max(1, min(args.steps, 2))
            for _ in range(reps):
                e2e_step()
            barrier()
            dt = (time.perf_counter() - t0) / reps
            tt = torch.tensor([dt], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e2e = dict(value=frames_step / float(tt.item()) / 1e6, unit=UNIT, h2d_bytes_per_step=K * N * fb,
                       d2h_bytes_per_step=K * N * ob, ms_per_step=float(tt.item()) * 1e3, steps=reps,
                       how="icw_session_process_host on pinned host buffers; H2D + kernels + D2H inside the timed region",
                       host_affinity=numa_note)
            del h_in, h_out
        except RuntimeError as ex:
            e2e = dict(value=None, unit=UNIT, error=str(ex)[:200])

    # ---- CPU baseline on this box's host cores (rank 0, N = 1 only) ------------------------------------
    cpu = None
    if wl.get("time_sharded"):
        args.no_cpu = True
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = host_cores()
        fr = min(N, 480_000)
        spc = max(1, int(12.0e6 // fr))               # ~12 s of single-core work per process
        f, busy, wall, kind = cpu_run(spec, fr, spc, cores)
        cpu = dict(value=f / busy / 1e6, unit=UNIT, cores=cores, kind=kind, cpu=cpu_model(),
                   sample=f"{cores} processes x {spc} fresh streams x {fr} frames of the same chain ({busy:.1f} s busy)")

    if rank == 0:
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
                    data="synthetic",
                    parity=dict(hilbert=wl["hilbert"],
                                note=("exact: bit-exact with the reference (tests/test_gpu_parity.py)" if wl["hilbert"] != "scan" else
                                      "scan: analytic signal within 1e-12 of the binary128 evaluation of the reference's filter; "
                                      "everything downstream byte-exact; the reference's own FP64 rounding noise (5e-5 of RMS for the "
                                      "default design) separates its PCM from ours (tests/test_gpu_scan.py)")),
                    config=dict(workload=wl["desc"], streams_per_gpu=K, frames_per_stream=N, frames_per_launch=units_per_launch,
                                hilbert=wl["hilbert"], l2="inputs larger than L2 (per-step input %.1f GB)" % (K * N * fb / 1e9)),
                    roofline=roofline, cpu_baseline=cpu, e2e=e2e, gpu_launches=int(launches), clocks=clocks,
                    kernel_ms={k: v["ms"] / args.steps for k, v in prof.items()})
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
