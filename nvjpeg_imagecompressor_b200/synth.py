"""Synthetic benchmark image of SURVEY.md Appendix B, generated on the device with torch integer ops
(bit-identical to the numpy/C generator used by the tests: integer-only hash + triangle waves)."""
import torch

_M = 0xFFFFFFFF


def _hash32(x, y, c, s):
    h = ((x * 0x9E3779B1) & _M) ^ ((y * 0x85EBCA77) & _M) ^ ((c * 0xC2B2AE3D) & _M) ^ ((s * 0x27D4EB2F) & _M)
    h = h ^ (h >> 15)
    h = (h * 0x2C1B3C6D) & _M
    h = h ^ (h >> 12)
    h = (h * 0x297A2D39) & _M
    h = h ^ (h >> 15)
    return h


def _tri(v, P):
    t = torch.remainder(v, 2 * P) - P
    return torch.div(t.abs() * 255, P, rounding_mode="floor")


def synth_rows(W, H, y0, rows, seed=0, amp=8, device="cuda"):
    """rows [y0, y0+rows) of the W x H image, uint8 [rows, W, 3] (B, G, R)."""
    x = torch.arange(W, device=device, dtype=torch.int64)[None, :, None]
    y = torch.arange(y0, y0 + rows, device=device, dtype=torch.int64)[:, None, None]
    c = torch.arange(3, device=device, dtype=torch.int64)[None, None, :]
    cell = (_hash32(x >> 5, y >> 5, torch.full_like(x, 7), seed) & 63) - 32
    base = torch.div(_tri(x + 3 * y + 37 * c, 512) + _tri(5 * x - 2 * y + 91 * c, 160) + _tri(y + 11 * c + 0 * x, 2048), 3,
                     rounding_mode="floor")
    noise = torch.remainder(_hash32(x, y, c, seed), 2 * amp + 1) - amp
    return (base + cell + noise).clamp_(0, 255).to(torch.uint8)


def synth(W, H, seed=0, amp=8, device="cuda", chunk=500):
    out = torch.empty((H, W, 3), dtype=torch.uint8, device=device)
    for y0 in range(0, H, chunk):
        r = min(chunk, H - y0)
        out[y0:y0 + r] = synth_rows(W, H, y0, r, seed, amp, device)
    return out
