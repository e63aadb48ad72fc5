"""Multi-GPU host logic on CPU: world_size-2 gloo process groups, the oracle standing in for the
kernels (allowed here: tests may use oracle/ as compute).  Checks the partitioning, the closed-form
segment state (oscillator counter incl. its wrap, mixer phase, dither offset, file position), the
nearest-neighbour hand-off and the counter reduction of in_cwave_b200.dist."""
import ctypes as C
import os
import tempfile

import numpy as np
import pytest

from in_cwave_b200 import dist as D
from in_cwave_b200 import spec as S
from in_cwave_b200 import synth


def test_shard_streams_partition():
    for n, w in ((4096, 8), (10, 4), (3, 8), (1, 1)):
        spans = [D.shard_streams(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def test_shard_time_partition():
    for n, w in ((33_177_600_000, 8), (1001, 2), (7, 4)):
        spans = [D.shard_time(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        assert all(a % 4 == 0 for a, _ in spans)


def test_closed_form_state_wraps_like_the_reference(oracle):
    spec = S.default_spec(fmt="cw_i16", sample_rate=8, need24bits=0, render_type=2)
    n = 8 * 1000 + 123
    raw = np.random.default_rng(0).integers(0, 256, size=n * S.frame_bytes(spec), dtype=np.uint8)
    st = oracle.port_process(spec, raw)["state"]
    cf = D.closed_form_state(spec, n)
    assert (cf.n_frame, cf.pos, cf.mt_drawn) == (int(st.n_frame), int(st.pos), int(st.mt[0].drawn))
    assert cf.n_frame == 123


class OracleBackend:
    """Stand-in for CudaBackend: same interface, the port oracle as the compute."""

    def __init__(self, oracle, spec):
        self.o, self.spec = oracle, spec
        self.st = oracle.new_state()

    def hilbert_state_after(self, raw_tail, quad0):
        st = self.o.new_state()
        st.quad[0] = st.quad[1] = quad0
        self.o.port_process(self.spec, np.asarray(raw_tail), state=st)
        hb = np.zeros((2, 2, 20))
        for c in range(2):
            for f in range(2):
                z, ix = st.lpf[c][f].z, st.lpf[c][f].ix
                hb[c, f, :] = [z[i] for i in range(20)]
                hb[c, f, 19] = ix          # the oracle's delay line is circular: ship its index along
        return hb.reshape(-1)

    def start_at(self, start, hb):
        st = self.st
        st.n_frame, st.pos = start.n_frame, start.pos
        st.quad[0] = st.quad[1] = start.quad
        for ch in range(2):
            for _ in range(start.mt_drawn):
                self.o.port().icwo_mt_u32(C.byref(st.mt[ch]))
        hb = np.asarray(hb).reshape(2, 2, 20)
        if np.any(hb):
            for c in range(2):
                for f in range(2):
                    for i in range(19):
                        st.lpf[c][f].z[i] = hb[c, f, i]
                    st.lpf[c][f].ix = int(hb[c, f, 19])

    def process(self, raw, d_out=None):
        return self.o.port_process(self.spec, np.asarray(raw), state=self.st)["pcm"]

    def counters(self):
        lin = [0.0 if p <= -555.0 else 10.0 ** (p / 20.0) for p in self.st.peak_db]
        return [int(self.st.clips[0]), int(self.st.clips[1])], lin


def _worker(rank, world, port, spec, n, level, outdir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyoracle
    fb = S.frame_bytes(spec)
    raw = synth.stream_bytes(spec, n, stream_id=77, level=level)
    a, b = D.shard_time(n, rank, world)
    be = OracleBackend(pyoracle, spec)
    pcm, clips, peaks, _ = D.run_time_sharded(be, dist, spec, raw[a * fb:b * fb], a, rank, world, warmup=1 << 30)
    np.save(os.path.join(outdir, f"pcm{rank}.npy"), pcm)
    np.save(os.path.join(outdir, f"cnt{rank}.npy"), np.array(clips + peaks, dtype=np.float64))
    dist.barrier()
    dist.destroy_process_group()


def _run_world(spec, n, level=0.25, world=2):
    import torch.multiprocessing as mp
    with tempfile.TemporaryDirectory() as td:
        port = 29500 + os.getpid() % 2000
        mp.spawn(_worker, args=(world, port, spec, n, level, td), nprocs=world, join=True)
        pcm = np.concatenate([np.load(os.path.join(td, f"pcm{r}.npy")) for r in range(world)])
        cnt = [np.load(os.path.join(td, f"cnt{r}.npy")) for r in range(world)]
    return pcm, cnt


def test_time_shards_closed_form_state_gloo(oracle):
    """Complex input (no filter state): oscillator counter, dither offset and file position are all
    the shards need -> the stitched PCM equals the single-run PCM byte for byte; counters reduced."""
    spec = S.config_c3(render_type=2, sample_rate=8000)     # TPDF dither, counter wraps inside the run
    n = 8000 * 1000 // 400 + 12345                          # 32345 frames
    pcm, cnt = _run_world(spec, n, level=2.5)
    ref = oracle.port_process(spec, synth.stream_bytes(spec, n, stream_id=77, level=2.5))
    assert np.array_equal(pcm, ref["pcm"])
    assert np.array_equal(cnt[0], cnt[1])                   # every rank holds the reduced counters
    st = ref["state"]
    assert [int(cnt[0][0]), int(cnt[0][1])] == [st.clips[0], st.clips[1]] and st.clips[0] > 0
    for c in range(2):
        assert abs(20 * np.log10(cnt[0][2 + c]) - st.peak_db[c]) < 1e-9


def test_time_shards_filter_state_handoff_gloo(oracle):
    """Real input: rank 0 ships its end-of-segment filter state to rank 1.  With the hand-off covering
    the whole first segment the stand-in's state is the true one, so the stitch is exact; this checks
    the plumbing (who sends what to whom, and that it is applied), not filter memory."""
    spec = S.config_c2(sample_rate=8000)
    n = 6000
    pcm, _ = _run_world(spec, n)
    ref = oracle.port_process(spec, synth.stream_bytes(spec, n, stream_id=77))
    assert np.array_equal(pcm, ref["pcm"])


def test_feedback_lists_are_not_time_sharded():
    """A list that reads a plug before it is written carries the bus from frame to frame: no closed form at a cut."""
    from in_cwave_b200 import dist as D, spec as S
    from test_oracle_vs_ref import FEEDBACK_NODES
    assert D.graph_has_feedback(S.default_spec(nodes=FEEDBACK_NODES))
    assert not D.graph_has_feedback(S.config_c3()) and not D.graph_has_feedback(S.config_c2())
    assert not D.graph_has_feedback(dict(S.default_spec(nodes=FEEDBACK_NODES), bypass=1))
