"""Multi-GPU partitioning of the chain: one process per GPU, torch.distributed for the plumbing.

Two ways the path shards:

* by STREAM -- independent files/channels: rank r owns streams [lo, hi); no communication at all
  (``shard_streams``).  This is what bench.py does by default ("scaling": "weak").
* by TIME inside one long stream -- rank r owns frames [a_r, b_r).  Everything a frame needs except
  the Hilbert filter state is a closed form of its index (oscillator counter, fs/4 mixer phase,
  dither generator offset, file position), so the only exchange is ONE nearest-neighbour hand-off:
  rank r-1 computes the filter state at b_{r-1} = a_r by running the converter over its last
  ``warmup`` frames from zero (the filters forget: |pole|^warmup is far below rounding) and sends
  those <= 1 KB to rank r (``run_time_sharded``); clip counters and peaks are all-reduced at the end.

The compute is behind a small backend interface so that the same logic runs on NCCL with the CUDA
session (``CudaBackend``) and, in the CPU tests, on gloo with a stand-in backend.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _abi, spec as _spec

WARMUP_FRAMES = 1 << 19     # 0.9997^(2^19) ~ 1e-68 for the slowest pole of the six designs
STATE_DOUBLES = 2 * 2 * _abi.MAX_ORD


def shard_streams(n_streams: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced [lo, hi) of streams for this rank."""
    base, extra = divmod(n_streams, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_time(n_frames: int, rank: int, world: int, align: int = 4) -> tuple[int, int]:
    """[a, b) of frames for this rank; interior boundaries are multiples of ``align`` (the mixer period)."""
    def cut(r):
        if r <= 0:
            return 0
        if r >= world:
            return n_frames
        return (n_frames * r // world) // align * align
    return cut(rank), cut(rank + 1)


@dataclass
class SegmentStart:
    """Closed-form part of a stream's state at absolute frame ``a`` of a file played from a fresh context."""
    n_frame: int
    pos: int
    quad: int
    mt_drawn: int


def closed_form_state(spec: dict, a: int, base: _abi.StreamState | None = None) -> SegmentStart:
    """State scalars at frame ``a`` (reference src/adv_modulator.c:611-625, src/lpf_hilbert_quad.c:155,
    src/sound_render.c:711-751).  ``base`` = the stream's state at frame 0 (default: fresh)."""
    wps = (0, 2, 4, 2, 24)[int(spec.get("render_type", 0))]
    n0 = int(base.n_frame) if base is not None else 0
    q0 = int(base.quad[0]) if base is not None else 0
    d0 = int(base.mt_drawn[0]) if base is not None else 0
    p0 = int(base.pos) if base is not None else 0
    if int(spec.get("is_frmod_scaled", 1)):
        n_frame = (n0 + a) % (int(spec.get("sample_rate", 48000)) * 1000)
    else:
        n_frame = n0 + a
    return SegmentStart(n_frame=n_frame, pos=p0 + a, quad=(q0 + a) & 3, mt_drawn=d0 + a * wps)


class CudaBackend:
    """The CUDA session as the compute engine of a time shard (scan mode)."""

    def __init__(self, engine, spec: dict):
        if spec.get("hilbert_mode", "exact") not in ("scan", 1):
            raise ValueError("time sharding needs hilbert_mode='scan': the exact recurrences are serial in time")
        if int(spec.get("nshape_type", 0)):
            raise ValueError("time sharding needs FLAT noise shaping: the error feedback is serial in time (SURVEY.md 8e)")
        self.engine, self.spec = engine, dict(spec)
        self.ses = engine.session(spec, 1)

    def hilbert_state_after(self, raw_tail, quad0: int) -> np.ndarray:
        """Filter state after running the converter over raw_tail (uint8 bytes, host or CUDA tensor) from zero."""
        import torch
        ses = self.engine.session(self.spec, 1)
        st = ses.get_state(0)
        st.quad[0] = st.quad[1] = quad0 & 3
        ses.set_state(0, st)
        fb = ses.frame_bytes
        if isinstance(raw_tail, np.ndarray):
            n = raw_tail.size // fb
            ses.process_host(raw_tail)
        else:
            n = raw_tail.numel() // fb
            out = torch.empty(n * ses.out_frame_bytes + 16, dtype=torch.uint8, device=raw_tail.device)
            ses.process_device(raw_tail, n, out, stream=torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
        st = ses.get_state(0)
        hb = np.array([[[st.hb[c][f][i] for i in range(_abi.MAX_ORD)] for f in range(2)] for c in range(2)])
        ses.close()
        return hb.reshape(-1)

    def start_at(self, start: SegmentStart, hb: np.ndarray) -> None:
        st = self.ses.get_state(0)
        st.n_frame, st.pos = start.n_frame, start.pos
        st.quad[0] = st.quad[1] = start.quad
        st.mt_drawn[0] = st.mt_drawn[1] = start.mt_drawn
        hb = np.asarray(hb, dtype=np.float64).reshape(2, 2, _abi.MAX_ORD)
        for c in range(2):
            for f in range(2):
                for i in range(_abi.MAX_ORD):
                    st.hb[c][f][i] = float(hb[c, f, i])
        st.hb_basis = 1
        self.ses.set_state(0, st)

    def process(self, raw, d_out=None):
        if isinstance(raw, np.ndarray):
            return self.ses.process_host(raw)[0]
        import torch
        n = raw.numel() // self.ses.frame_bytes
        self.ses.process_device(raw, n, d_out, stream=torch.cuda.current_stream().cuda_stream)
        return d_out

    def counters(self) -> tuple[list[int], list[float]]:
        st = self.ses.get_state(0)
        return [int(st.clips[0]), int(st.clips[1])], [float(st.peak[0]), float(st.peak[1])]


def handoff(dist, state_out: np.ndarray | None, rank: int, world: int, device="cpu", group=None) -> np.ndarray:
    """The one collective of the time-sharded path: every rank but the last sends its end-of-segment
    filter state to its right neighbour; rank 0 starts from zeros.  <= 1 KB per stream."""
    import torch
    recv = torch.zeros(STATE_DOUBLES, dtype=torch.float64, device=device)
    ops = []
    if rank + 1 < world:
        send = torch.as_tensor(np.asarray(state_out, dtype=np.float64), device=device).contiguous()
        ops.append(dist.P2POp(dist.isend, send, rank + 1, group=group))
    if rank > 0:
        ops.append(dist.P2POp(dist.irecv, recv, rank - 1, group=group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return recv.cpu().numpy()


def reduce_counters(dist, clips: list[int], peaks: list[float], device="cpu", group=None):
    """Sum of clip counters, max of peaks over the shards (reference: one global pair of accumulators,
    src/adv_modulator.c:54-55)."""
    import torch
    c = torch.tensor(clips, dtype=torch.int64, device=device)
    p = torch.tensor(peaks, dtype=torch.float64, device=device)
    dist.all_reduce(c, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(p, op=dist.ReduceOp.MAX, group=group)
    return [int(v) for v in c.cpu()], [float(v) for v in p.cpu()]


def run_time_sharded(backend, dist, spec: dict, my_raw, a: int, rank: int, world: int, device="cpu",
                     warmup: int = WARMUP_FRAMES, d_out=None, group=None):
    """Process this rank's frames [a, a + len(my_raw)/frame_bytes) of one long stream.

    my_raw: this rank's input bytes (numpy uint8 or CUDA uint8 tensor).  Returns (pcm, clips, peaks)
    with the counters already reduced over all ranks."""
    fb = _spec.frame_bytes(spec)
    n = (my_raw.size if isinstance(my_raw, np.ndarray) else my_raw.numel()) // fb
    # 1. state at my segment end, from a warm-up over my own last frames -> right neighbour
    state_out = None
    if rank + 1 < world:
        w = min(warmup, n)
        tail = my_raw[(n - w) * fb: n * fb]
        state_out = backend.hilbert_state_after(tail, (closed_form_state(spec, a + n - w).quad))
    state_in = handoff(dist, state_out, rank, world, device=device, group=group)
    # 2. my segment from the received state and the closed-form scalars
    backend.start_at(closed_form_state(spec, a), state_in)
    pcm = backend.process(my_raw, d_out)
    clips, peaks = backend.counters()
    clips, peaks = reduce_counters(dist, clips, peaks, device=device, group=group)
    return pcm, clips, peaks
