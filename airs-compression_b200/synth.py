"""Synthetic telemetry-like sample generators (integer only).

numpy implementation for the host side; the same arithmetic is used by the
torch version in batch.py so that CPU oracle and GPU see identical inputs.
x[c, i] = (walk[c, i] + noise[c, i]) mod 2^16, walk = base + cumsum(step in [-3, 3]),
noise in [-15, 15]; `esc` of 256 samples get 16 random bits XORed in (escape-heavy data).
"""
import numpy as np

_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
_G = np.uint64(0x9E3779B97F4A7C15)


def mix(z):
    """splitmix64 finaliser on uint64 arrays."""
    with np.errstate(over="ignore"):
        z = (np.asarray(z, dtype=np.uint64) + _G)
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        return z ^ (z >> np.uint64(31))


def chunks(seed, first_chunk, n_chunks, n_samples, esc=0):
    """uint16 array [n_chunks, n_samples] of independent random-walk chunks."""
    with np.errstate(over="ignore"):
        c = np.arange(first_chunk, first_chunk + n_chunks, dtype=np.uint64)
        hc = mix(np.uint64(seed) + c)
        i = np.arange(n_samples, dtype=np.uint64)
        h = mix(hc[:, None] + i[None, :])
    step = (h & np.uint64(3)).astype(np.int64) - ((h >> np.uint64(2)) & np.uint64(3)).astype(np.int64)
    noise = ((h >> np.uint64(8)) & np.uint64(15)).astype(np.int64) - \
        ((h >> np.uint64(12)) & np.uint64(15)).astype(np.int64)
    base = 0x8000 + ((hc >> np.uint64(48)) & np.uint64(0x3FFF)).astype(np.int64)
    walk = base[:, None] + np.cumsum(step, axis=1)
    x = (walk + noise) & 0xFFFF
    if esc:
        hit = ((h >> np.uint64(32)) & np.uint64(0xFF)).astype(np.int64) < esc
        x = np.where(hit, x ^ ((h >> np.uint64(40)) & np.uint64(0xFFFF)).astype(np.int64), x)
    return x.astype(np.uint16)


def frames(seed, context, n_frames, n_samples, drift=3):
    """uint16 [n_frames, n_samples]: a fixed scene (random walk) + slow drift + fresh noise."""
    scene = chunks(seed, context, 1, n_samples)[0].astype(np.int64)
    with np.errstate(over="ignore"):
        k = np.arange(n_frames, dtype=np.uint64)
        hk = mix(mix(np.uint64(seed) + np.uint64(context)) ^ (k + np.uint64(0x5151)))
        i = np.arange(n_samples, dtype=np.uint64)
        h = mix(hk[:, None] + i[None, :])
    noise = ((h >> np.uint64(8)) & np.uint64(15)).astype(np.int64) - \
        ((h >> np.uint64(12)) & np.uint64(15)).astype(np.int64)
    kk = np.arange(n_frames, dtype=np.int64)
    x = (scene[None, :] + ((kk * drift) >> 4)[:, None] + noise) & 0xFFFF
    return x.astype(np.uint16)


# ---------------------------------------------------------------------------
# torch versions (same integers; int64 two's-complement arithmetic == uint64 mod 2^64)
# ---------------------------------------------------------------------------

def _t_mix(z):
    import torch
    def lsr(v, k):
        return (v >> k) & ((1 << (64 - k)) - 1)
    z = z + torch.tensor(-7046029254386353131, dtype=torch.int64, device=z.device)  # 0x9E3779B97F4A7C15
    z = (z ^ lsr(z, 30)) * torch.tensor(-4658895280553007687, dtype=torch.int64, device=z.device)  # 0xBF58476D1CE4E5B9
    z = (z ^ lsr(z, 27)) * torch.tensor(-7723592293110705685, dtype=torch.int64, device=z.device)  # 0x94D049BB133111EB
    return z ^ lsr(z, 31)


def _t_noise(h):
    return ((h >> 8) & 15) - ((h >> 12) & 15)


def chunks_torch(seed, first_chunk, n_chunks, n_samples, esc=0, device="cuda", slab=1 << 26):
    """torch twin of chunks(): returns a uint16-valued int16 tensor [n_chunks, n_samples] on device."""
    import torch
    out = torch.empty((n_chunks, n_samples), dtype=torch.int16, device=device)
    per = max(1, slab // n_samples)
    i = torch.arange(n_samples, dtype=torch.int64, device=device)
    for c0 in range(0, n_chunks, per):
        c1 = min(n_chunks, c0 + per)
        c = torch.arange(first_chunk + c0, first_chunk + c1, dtype=torch.int64, device=device)
        hc = _t_mix(c + seed)
        h = _t_mix(hc[:, None] + i[None, :])
        step = (h & 3) - ((h >> 2) & 3)
        base = 0x8000 + ((hc >> 48) & 0x3FFF)
        x = base[:, None] + torch.cumsum(step, dim=1) + _t_noise(h)
        if esc:
            hit = ((h >> 32) & 0xFF) < esc
            x = torch.where(hit, (x & 0xFFFF) ^ ((h >> 40) & 0xFFFF), x)
        out[c0:c1] = (x & 0xFFFF).to(torch.int32).to(torch.int16)
        del h, step, x
    return out


def frames_torch(seed, first_context, n_contexts, n_frames, n_samples, drift=3, device="cuda"):
    """torch twin of frames() for many contexts: int16 tensor [n_contexts, n_frames, n_samples]."""
    import torch
    out = torch.empty((n_contexts, n_frames, n_samples), dtype=torch.int16, device=device)
    i = torch.arange(n_samples, dtype=torch.int64, device=device)
    k = torch.arange(n_frames, dtype=torch.int64, device=device)
    for ci in range(n_contexts):
        ctx = first_context + ci
        scene = chunks_torch(seed, ctx, 1, n_samples, device=device)[0].to(torch.int64) & 0xFFFF
        base = _t_mix(torch.tensor([seed + ctx], dtype=torch.int64, device=device))
        hk = _t_mix(base ^ (k + 0x5151))
        h = _t_mix(hk[:, None] + i[None, :])
        x = scene[None, :] + ((k * drift) >> 4)[:, None] + _t_noise(h)
        out[ci] = (x & 0xFFFF).to(torch.int32).to(torch.int16)
    return out
