"""Synthetic inputs of the benchmark shapes: deterministic white noise plus
two sines (997 Hz and 0.31*fs/2), total level about -12 dBFS so that clips stay rare.

numpy on the host for tests and e2e; ``device_*`` variants fill torch CUDA tensors directly for the
large resident workloads (there is no dataset to load -- the bench says "data": "synthetic").
"""
from __future__ import annotations

import numpy as np


def _mix(stream_id: int) -> int:
    z = (0x1C0A7E5EED ^ (stream_id * 0x9E3779B97F4A7C15)) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return (z ^ (z >> 31)) & 0xFFFFFFFF


def signal(n: int, n_ch: int, sample_rate: int, stream_id: int = 0, level: float = 0.25) -> np.ndarray:
    """float64 [n, n_ch] in [-1, 1): noise (level/2 peak) + two sines (level/4 each)."""
    rng = np.random.default_rng(_mix(stream_id))
    t = np.arange(n, dtype=np.float64)[:, None]
    ph = rng.random((1, n_ch)) * 2 * np.pi
    x = (rng.random((n, n_ch)) - 0.5) * level
    x += 0.25 * level * np.sin(2 * np.pi * 997.0 / sample_rate * t + ph)
    x += 0.25 * level * np.sin(2 * np.pi * 0.155 * t + 2 * ph)
    return x


def encode(x: np.ndarray, fmt: str) -> np.ndarray:
    """[-1,1) float64 [n, cols] -> raw little-endian sample bytes (uint8, 1-D) of a WAV/CWAVE format.
    For complex formats the columns are (L.I, L.Q[, R.I, R.Q]) and are scaled to +-32768 units."""
    if fmt == "wav_f32":
        return np.ascontiguousarray(x.astype("<f4")).view(np.uint8).ravel()
    if fmt == "wav_i16":
        return np.ascontiguousarray(np.clip(np.round(x * 32768.0), -32768, 32767).astype("<i2")).view(np.uint8).ravel()
    if fmt == "wav_i32":
        return np.ascontiguousarray(np.clip(np.round(x * 2147483648.0), -2**31, 2**31 - 1).astype("<i4")).view(np.uint8).ravel()
    if fmt == "wav_u8":
        return np.ascontiguousarray((np.clip(np.round(x * 128.0), -128, 127) + 128).astype(np.uint8)).ravel()
    if fmt == "wav_i24":
        v = np.clip(np.round(x * 8388608.0), -2**23, 2**23 - 1).astype("<i4")
        b = np.ascontiguousarray(v).view(np.uint8).reshape(-1, 4)[:, :3]
        return np.ascontiguousarray(b).ravel()
    if fmt == "cw_f32":
        return np.ascontiguousarray((x * 32768.0).astype("<f4")).view(np.uint8).ravel()
    if fmt == "cw_f64":
        return np.ascontiguousarray((x * 32768.0).astype("<f8")).view(np.uint8).ravel()
    if fmt == "cw_i16":
        return np.ascontiguousarray(np.clip(np.round(x * 32768.0), -32768, 32767).astype("<i2")).view(np.uint8).ravel()
    if fmt == "cw_i16f32":
        n, cols = x.shape
        out = np.zeros((n, cols // 2, 6), dtype=np.uint8)
        i16 = np.clip(np.round(x[:, 0::2] * 32768.0), -32768, 32767).astype("<i2")
        f32 = (x[:, 1::2] * 32768.0).astype("<f4")
        out[:, :, 0:2] = np.ascontiguousarray(i16).view(np.uint8).reshape(n, cols // 2, 2)
        out[:, :, 2:6] = np.ascontiguousarray(f32).view(np.uint8).reshape(n, cols // 2, 4)
        return out.ravel()
    raise ValueError(fmt)


def stream_bytes(spec: dict, n: int, stream_id: int = 0, level: float = 0.25) -> np.ndarray:
    """Raw input bytes of one synthetic stream for a chain spec."""
    fmt = spec.get("fmt", "wav_f32")
    nch = int(spec.get("n_channels", 2))
    cols = nch * (2 if fmt.startswith("cw_") else 1)
    return encode(signal(n, cols, int(spec.get("sample_rate", 48000)), stream_id, level), fmt)


def device_fill(spec: dict, n_streams: int, n: int, device, row_align: int = 16):
    """Resident synthetic input: torch uint8 [n_streams, stride] generated on the GPU (noise + two
    sines, same family as ``signal``), in blocks of at most 2^24 sample values at a time."""
    import torch
    from .spec import frame_bytes
    fmt = spec.get("fmt", "wav_f32")
    nch = int(spec.get("n_channels", 2))
    cols = nch * (2 if fmt.startswith("cw_") else 1)
    sr = int(spec.get("sample_rate", 48000))
    fb = frame_bytes(spec)
    stride = (n * fb + row_align - 1) // row_align * row_align
    out = torch.zeros((n_streams, stride), dtype=torch.uint8, device=device)
    g = torch.Generator(device=device)
    g.manual_seed(0x1C0A7E5EED)
    tblk = max(1, min(n, (1 << 24) // cols))
    sblk = max(1, min(n_streams, (1 << 24) // (tblk * cols)))
    for t0 in range(0, n, tblk):
        tn = min(tblk, n - t0)
        t = torch.arange(t0, t0 + tn, dtype=torch.float64, device=device)[:, None]
        tone = (0.0625 * torch.sin(2 * torch.pi * 997.0 / sr * t) + 0.0625 * torch.sin(2 * torch.pi * 0.155 * t)).to(torch.float32)
        for s0 in range(0, n_streams, sblk):
            k = min(sblk, n_streams - s0)
            x = (torch.rand((k, tn, cols), generator=g, device=device, dtype=torch.float32) - 0.5) * 0.25 + tone[None]
            if fmt == "wav_f32":
                b = x.contiguous().view(torch.uint8).reshape(k, -1)
            elif fmt == "cw_f32":
                b = (x * 32768.0).contiguous().view(torch.uint8).reshape(k, -1)
            elif fmt == "wav_i16":
                b = torch.clamp(torch.round(x * 32768.0), -32768, 32767).to(torch.int16).contiguous().view(torch.uint8).reshape(k, -1)
            elif fmt == "wav_i24":
                v = torch.clamp(torch.round(x.double() * 8388608.0), -2**23, 2**23 - 1).to(torch.int32)
                b = v.contiguous().view(torch.uint8).reshape(k, tn, cols, 4)[..., :3].contiguous().reshape(k, -1)
            else:
                raise ValueError(f"device_fill: format {fmt} not wired")
            out[s0:s0 + k, t0 * fb:(t0 + tn) * fb] = b
    return out
