"""Two real GPUs, two processes, NCCL: the time-sharded stream through the C entry points
(icw_comm_init / icw_session_seek_closed_form / icw_session_handoff / icw_session_reduce_counters) stitches to the
bytes of a one-GPU run.  Skipped on a one-GPU box (NCCL refuses two ranks on one device); bench.py --gpus N runs the
same path with its own stitch check."""
import os
import tempfile

import numpy as np
import pytest

from in_cwave_b200 import spec as S
from in_cwave_b200 import synth

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, spec, n, outdir):
    import torch
    import torch.distributed as dist
    import in_cwave_b200 as icw
    from in_cwave_b200 import dist as D
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device(f"cuda:{rank}")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    fb, ob = S.frame_bytes(spec), S.out_frame_bytes(spec)
    raw = synth.stream_bytes(spec, n, stream_id=91)
    a, b = D.shard_time(n, rank, world)
    eng = icw.Engine(rank)
    comm = D.make_comm(dist, rank, world, dev)
    be = D.CudaBackend(eng, spec, comm=comm)
    mine = torch.from_numpy(raw[a * fb:b * fb].copy()).to(dev)
    out = torch.empty((b - a) * ob + 16, dtype=torch.uint8, device=dev)
    _, clips, peaks, ms = D.run_time_sharded(be, dist, spec, mine, a, rank, world, device=dev, d_out=out, warmup=1 << 19)
    torch.cuda.synchronize()
    np.save(os.path.join(outdir, f"pcm{rank}.npy"), out[: (b - a) * ob].cpu().numpy())
    np.save(os.path.join(outdir, f"cnt{rank}.npy"), np.array(clips + peaks + [ms], dtype=np.float64))
    dist.barrier()
    be.close()
    D.free_comm(comm)
    dist.destroy_process_group()


@pytest.mark.parametrize("render_type", [2, 3])
def test_two_rank_nccl_handoff_stitches_to_the_one_gpu_run(engine, render_type):
    import torch
    import torch.multiprocessing as mp
    from util import pcm_report
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    spec = S.config_c2(hilbert_mode="scan", render_type=render_type)      # TPDF; sloped TPDF carries a draw across the cut
    n = 2 * (1 << 19) + 40000
    raw = synth.stream_bytes(spec, n, stream_id=91)
    whole = engine.session(spec, 1).process_host(raw)[0]
    with tempfile.TemporaryDirectory() as td:
        port = 29500 + os.getpid() % 2000
        mp.spawn(_worker, args=(2, port, spec, n, td), nprocs=2, join=True)
        pcm = np.concatenate([np.load(os.path.join(td, f"pcm{r}.npy")) for r in range(2)])
        cnt = [np.load(os.path.join(td, f"cnt{r}.npy")) for r in range(2)]
    rep = pcm_report(pcm, whole, 3)
    print(f"[2-rank NCCL stitch, render_type {render_type}] {rep}, hand-off {cnt[0][4]:.3f} / {cnt[1][4]:.3f} ms")
    assert rep["max_lsb"] <= 1 and rep["mismatches"] <= 2
    assert np.array_equal(cnt[0][:4], cnt[1][:4])           # both ranks hold the reduced counters
