#!/usr/bin/env python3
"""N1 end to end (reference src/xwave_reader.c:838-904, src/transcode.c:40-118): one large 24-bit WAV in tmpfs
through libicw_plugin.so's transcode entry points in 64 KB getData calls -- file read on the reader thread, H2D, kernels,
D2H -- against (a) reading the file alone and (b) the GPU part alone, and the compiled reference (oracle/_ref, one core:
its transcode is one thread) on a slice of the same file.  Prints one JSON object; run on the GPU box.

    python tools/transcode_bigfile.py [--gib 1.9] [--dir /dev/shm] [--ref-frames 6000000]
"""
import argparse
import ctypes as C
import json
import os
import struct
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from in_cwave_b200 import plugin, spec as S, synth    # noqa: E402


def bind(L):
    ip = C.POINTER(C.c_int)
    L.winampGetExtendedRead_open.argtypes = [C.c_char_p, ip, ip, ip, ip]
    L.winampGetExtendedRead_open.restype = C.c_ssize_t
    L.winampGetExtendedRead_getData.argtypes = [C.c_ssize_t, C.c_void_p, C.c_int, ip]
    L.winampGetExtendedRead_getData.restype = C.c_ssize_t
    L.winampGetExtendedRead_close.argtypes = [C.c_ssize_t]
    L.winampGetExtendedRead_close.restype = None
    return L


def pull(L, path, chunk=65536):
    """open -> getData until 0 -> close; returns (bytes served, seconds, crc-like checksum of the first MiB)."""
    info = [C.c_int() for _ in range(4)]
    t0 = time.perf_counter()
    h = L.winampGetExtendedRead_open(str(path).encode(), *[C.byref(i) for i in info])
    if not h:
        raise SystemExit(f"open failed: {path}")
    buf = (C.c_char * chunk)()
    kill = C.c_int(0)
    total, head = 0, bytearray()
    while True:
        got = L.winampGetExtendedRead_getData(h, buf, chunk, C.byref(kill))
        if got <= 0:
            break
        if len(head) < (1 << 20):
            head += buf.raw[:got]
        total += got
    L.winampGetExtendedRead_close(h)
    return total, time.perf_counter() - t0, bytes(head[: 1 << 20])


def write_wav(path, spec, frames, block=1 << 22):
    fb = S.frame_bytes(spec)
    sr = spec["sample_rate"]
    data_len = frames * fb
    fmtc = struct.pack("<HHIIHH", 1, 2, sr, sr * fb, fb, 24)
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 4 + 8 + len(fmtc) + 8 + data_len) + b"WAVE" + b"fmt " + struct.pack("<I", len(fmtc)) + fmtc +
                b"data" + struct.pack("<I", data_len))
        one = np.asarray(synth.stream_bytes(spec, block, stream_id=5), dtype=np.uint8).tobytes()
        left = data_len
        while left > 0:
            n = min(left, len(one))
            f.write(one[:n])
            left -= n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gib", type=float, default=1.9)
    ap.add_argument("--dir", default="/dev/shm")
    ap.add_argument("--ref-frames", type=int, default=6_000_000)
    a = ap.parse_args()
    spec = S.config_c2(hilbert_mode="scan")                 # 192 kHz i24 stereo, shift + TPDF, 24-bit out (C2)
    fb = S.frame_bytes(spec)
    frames = int(a.gib * (1 << 30)) // fb
    big = Path(a.dir) / "icw_big.wav"
    small = Path(a.dir) / "icw_ref_slice.wav"
    write_wav(big, spec, frames)
    write_wav(small, spec, a.ref_frames)
    out = dict(file_bytes=os.path.getsize(big), frames=frames, seconds_of_audio=frames / spec["sample_rate"], getdata_bytes=65536)

    # (a) the read alone
    t0 = time.perf_counter()
    with open(big, "rb", buffering=0) as f:
        b = bytearray(1 << 24)
        while f.readinto(b):
            pass
    out["read_alone_s"] = time.perf_counter() - t0

    # ours
    L = bind(plugin.lib())
    plugin.lib().icwp_reset()
    plugin.configure(spec)
    pull(L, small)                                          # warm: engine, tables, page-locked blocks
    plugin.lib().icwp_reset()
    plugin.configure(spec)
    plugin.io_stats(reset=True)
    served, wall, head = pull(L, big)
    st = plugin.io_stats()
    out.update(served_bytes=served, wall_s=wall, gpu_s=st["gpu_s"], read_s=st["read_s"], wait_s=st["wait_s"],
               blocks_prefetched=st["blocks_prefetched"], blocks_sync=st["blocks_sync"],
               mframes_per_s=frames / wall / 1e6, wall_over_max_read_gpu=wall / max(out["read_alone_s"], st["gpu_s"]))

    # the compiled reference, one thread (its transcode path is serial), on the slice
    try:
        from oracle import pyoracle as po
        if po.have_ref():
            R = bind(po.ref())
            cfg = po.make_refcfg(dict(spec, hilbert_mode="exact"))
            po.ref().icwref_reset(C.byref(cfg))
            nodes = spec["nodes"]
            arr = (po.Node * len(nodes))()
            for i, nd in enumerate(nodes):
                po.fill_node(arr[i], nd)
            po.ref().icwref_set_graph(arr, len(nodes), 0)
            rs, rw, _ = pull(R, small)
            out.update(reference_slice_frames=a.ref_frames, reference_wall_s=rw, reference_mframes_per_s=a.ref_frames / rw / 1e6,
                       reference_wall_for_the_big_file_s=rw * frames / a.ref_frames, speedup_wall=(rw * frames / a.ref_frames) / wall)
    except Exception as ex:                                  # noqa: BLE001
        out["reference"] = f"unavailable: {ex}"
    os.remove(big)
    os.remove(small)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
