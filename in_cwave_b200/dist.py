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
    frames: int = 0         # the offset these were computed for


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
    return SegmentStart(n_frame=n_frame, pos=p0 + a, quad=(q0 + a) & 3, mt_drawn=d0 + a * wps, frames=a)


def make_comm(dist, rank: int, world: int, device="cpu"):
    """Our own NCCL communicator for the hand-off (include/icw_b200.h icw_comm_*): rank 0 makes the 128-byte id,
    torch.distributed -- the plumbing -- broadcasts it, every rank joins.  Returns an opaque handle (int) or None
    for world == 1."""
    import torch
    if world <= 1:
        return None
    L = _abi.lib()
    buf = C.create_string_buffer(_abi.COMM_ID_BYTES)
    if rank == 0:
        _abi.check(L.icw_comm_unique_id(buf))
    t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).to(device)
    dist.broadcast(t, src=0)
    idb = bytes(t.cpu().numpy().tobytes())
    comm = C.c_void_p()
    _abi.check(L.icw_comm_init(idb, rank, world, C.byref(comm)))
    return comm


def free_comm(comm) -> None:
    if comm:
        _abi.check(_abi.lib().icw_comm_destroy(comm))


def graph_has_feedback(spec: dict) -> bool:
    """The host's test (fold_spec, csrc/icw_api.cu): some node reads a plug that only a later node -- or the node itself
    -- writes in the same frame, i.e. the previous frame's value."""
    if spec.get("bypass"):
        return False
    nodes = spec["nodes"]
    written = {0}
    for i, nd in enumerate(nodes):
        later = {m["out"] for m in nodes[i:] if m["mode"] != "master"}
        if any(k in later and k not in written for k in nd.get("inputs", [0])):
            return True
        if nd["mode"] != "master":
            written.add(nd["out"])
    return False


class CudaBackend:
    """The CUDA session as the compute engine of a time shard (scan mode).

    With a communicator (``make_comm``) the whole step runs through the C ABI and the filter state never leaves
    device memory: icw_session_seek_closed_form + a warm-up run on a second session, icw_session_handoff
    (ncclSend / ncclRecv straight out of / into the sessions' device-resident state), the shard itself,
    icw_session_reduce_counters (``native_step``).  Without one (single GPU, CPU tests with a stand-in backend)
    ``run_time_sharded`` drives the generic numpy path below."""

    def __init__(self, engine, spec: dict, comm=None):
        if spec.get("hilbert_mode", "exact") not in ("scan", 1):
            raise ValueError("time sharding needs hilbert_mode='scan': the exact recurrences are serial in time")
        if int(spec.get("nshape_type", 0)):
            raise ValueError("time sharding needs FLAT noise shaping: the error feedback is serial in time (SURVEY.md 8e)")
        if graph_has_feedback(spec):
            raise ValueError("time sharding needs a feed-forward DSP list: a plug read before it is written carries the bus "
                             "from frame to frame (reference src/adv_modulator.c:634-751), which is serial in time")
        self.engine, self.spec = engine, dict(spec)
        self.ses = engine.session(spec, 1)
        self.comm = comm
        self._warm = None
        self.last_handoff_ms = None

    def close(self) -> None:
        if self._warm is not None:
            self._warm.close()
            self._warm = None
        self.ses.close()

    def warm_session(self):
        if self._warm is None:
            self._warm = self.engine.session(self.spec, 1)
        return self._warm

    # ---- the native path: everything through the C entry points ------------------------------------------
    def native_step(self, my_raw, a: int, rank: int, world: int, d_out, warmup: int = WARMUP_FRAMES):
        """Frames [a, a + n) of the long stream on this rank; returns (d_out, clips, peak_db, handoff_ms)."""
        import torch
        L = _abi.lib()
        cs = torch.cuda.current_stream().cuda_stream or 1
        fb = self.ses.frame_bytes
        n = my_raw.numel() // fb
        warm = None
        if rank + 1 < world:
            if n < warmup:
                raise ValueError(f"shard of {n} frames is shorter than the {warmup}-frame warm-up the hand-off state needs; "
                                 "use fewer ranks or pass the left neighbours' frames")
            warm = self.warm_session()
            warm.reset()
            _abi.check(L.icw_session_seek_closed_form(warm._h, 0, a + n - warmup, None))
            tail = my_raw[(n - warmup) * fb: n * fb]
            scratch = self._scratch(warmup * self.ses.out_frame_bytes + 16, my_raw.device)
            warm.process_device(tail, warmup, scratch, stream=cs)
        self.ses.reset()
        _abi.check(L.icw_session_seek_closed_form(self.ses._h, 0, a, None))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _abi.check(L.icw_session_handoff(warm._h if warm is not None else None, self.ses._h if rank > 0 else None,
                                         self.comm, rank, world, cs))
        e1.record()
        self.ses.process_device(my_raw, n, d_out, stream=cs)
        clips, peak = (C.c_uint64 * 2)(), (C.c_double * 2)()
        _abi.check(L.icw_session_reduce_counters(self.ses._h, self.comm, cs, C.byref(clips), C.byref(peak)))
        self.last_handoff_ms = e0.elapsed_time(e1)
        return d_out, [int(clips[0]), int(clips[1])], [float(peak[0]), float(peak[1])], self.last_handoff_ms

    def _scratch(self, nbytes: int, device):
        import torch
        if getattr(self, "_scr", None) is None or self._scr.numel() < nbytes:
            self._scr = torch.empty(nbytes, dtype=torch.uint8, device=device)
        return self._scr

    # ---- the generic path (single GPU emulation of the split; tests) ---------------------------------------
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
        L = _abi.lib()
        _abi.check(L.icw_session_seek_closed_form(self.ses._h, 0, start.frames, None))
        st = self.ses.get_state(0)
        assert (int(st.n_frame), int(st.pos), int(st.quad[0]), int(st.mt_drawn[0])) == \
            (start.n_frame, start.pos, start.quad, start.mt_drawn), "closed forms of dist.py and the C ABI disagree"
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

    my_raw: this rank's input bytes (numpy uint8 or CUDA uint8 tensor).  Returns (pcm, clips, peaks, handoff_ms):
    the counters already reduced over all ranks (peaks LINEAR on the generic path, dB on the native one -- see
    ``CudaBackend.native_step``), handoff_ms the device time of the collective or None."""
    if getattr(backend, "comm", None) is not None and not isinstance(my_raw, np.ndarray):
        return backend.native_step(my_raw, a, rank, world, d_out, warmup=min(warmup, WARMUP_FRAMES))
    fb = _spec.frame_bytes(spec)
    n = (my_raw.size if isinstance(my_raw, np.ndarray) else my_raw.numel()) // fb
    if int(spec.get("render_type", 0)) == 3 and rank > 0 and not hasattr(backend, "ses"):
        raise ValueError("sloped TPDF carries the previous sample's draw: only the CUDA backend regenerates it at a cut")
    # 1. state at my segment end, from a warm-up over my own last frames -> right neighbour
    state_out = None
    if rank + 1 < world:
        if n < warmup and warmup <= WARMUP_FRAMES:
            raise ValueError(f"shard of {n} frames is shorter than the {warmup}-frame warm-up the hand-off state needs")
        w = min(warmup, n)
        tail = my_raw[(n - w) * fb: n * fb]
        state_out = backend.hilbert_state_after(tail, (closed_form_state(spec, a + n - w).quad))
    state_in = handoff(dist, state_out, rank, world, device=device, group=group)
    # 2. my segment from the received state and the closed-form scalars
    backend.start_at(closed_form_state(spec, a), state_in)
    pcm = backend.process(my_raw, d_out)
    clips, peaks = backend.counters()
    clips, peaks = reduce_counters(dist, clips, peaks, device=device, group=group)
    return pcm, clips, peaks, None
