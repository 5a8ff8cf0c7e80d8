"""Host logic of the multi-GPU strip encoder (nvjpeg_imagecompressor_b200/strips.py) over gloo, world_size 2 and 3,
with a CPU backend built from the checker: row partition, DC-predictor exchange, histogram all-reduce, bit-phase
seams and the final gather must reproduce the single-stream JPEG byte for byte."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class OracleBackend:
    def __init__(self, O, quality, optimize, css):
        self.O, self.q, self.opt, self.css = O, quality, optimize, css
        self.hist = torch.zeros(4 * 257, dtype=torch.int32)
        self.last_dc = torch.zeros(4, dtype=torch.int16)
        self.pred_in = torch.zeros(4, dtype=torch.int16)
        self.strip_bits = torch.zeros(2, dtype=torch.int64)
        self.out_len = torch.zeros(1, dtype=torch.int64)
        self.out = np.zeros(0, np.uint8)

    def phase1(self, img, step, W, rows):
        O = self.O
        self.g = O.geometry(W, rows, self.css)
        self.coef = O.forward(img, self.css, self.q)
        last = self.coef.reshape(-1, self.g.bpm, 64)[-1]
        b = self.g.bpm
        self.last_dc[:3] = torch.tensor([int(last[b - 3, 0]), int(last[b - 2, 0]), int(last[b - 1, 0])], dtype=torch.int16)
        self.hist.zero_()

    def phase1b(self):
        if self.opt:
            h = self.O.histogram(self.coef, self.g.bpm, self.pred_in[:3].numpy())
            self.hist.copy_(torch.from_numpy(h.astype(np.int32).ravel()))

    def phase2(self, W, H):
        O = self.O
        if self.opt:
            h = self.hist.numpy().astype(np.uint32).reshape(4, 257)
            tb = [O.gen_optimal_table(h[t]) for t in range(4)]
            self.bits, self.vals = np.stack([t[0] for t in tb]), np.stack([t[1] for t in tb])
        else:
            self.bits, self.vals = O.std_tables()
        self.raw, nbits = O.entropy_bits(self.coef, self.g.bpm, self.bits, self.vals, self.pred_in[:3].numpy())
        first = int.from_bytes(bytes(self.raw[:4].tolist() + [0] * 4)[:4], "big")
        self.strip_bits[0], self.strip_bits[1] = nbits, first
        self.hdr = O.headers(W, H, self.css, O.quant_tables(self.q), self.bits, self.vals)

    def phase3(self, skip, ext, flags):
        T = int(self.strip_bits[0])
        bits = np.unpackbits(self.raw)[:T][skip:]
        nbytes = (len(bits) + 7) // 8
        tail = np.concatenate([np.unpackbits(np.array([ext], np.uint8)), np.ones(8, np.uint8)])
        bits = np.concatenate([bits, tail])[: nbytes * 8]
        body = self.O.stuff(np.packbits(bits), nbytes * 8)
        parts = ([self.hdr] if flags & 1 else []) + [body] + ([np.array([0xFF, 0xD9], np.uint8)] if flags & 2 else [])
        self.out = np.concatenate(parts)
        self.out_len[0] = self.out.size

    def out_view(self, n):
        return torch.from_numpy(self.out[:n].copy())


def _worker(rank, world, port, W, H, css, q, opt, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle as O
    from nvjpeg_imagecompressor_b200.strips import StripEncoder
    img = O.synth(W, H, 3, 8)
    enc = StripEncoder(W, H, q, opt, css, backend=OracleBackend(O, q, opt, css))
    strip = np.ascontiguousarray(img[enc.y0:enc.y1])
    enc.encode_strip(strip, W * 3)
    out = enc.gather_jpeg(0)
    if rank == 0:
        want = O.encode(img, css, q, opt)
        ret["ok"] = bool(np.array_equal(out.numpy(), want))
        ret["n"] = int(out.numel())
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,W,H,css,q,opt", [(2, 96, 80, 1, 95, 1), (2, 70, 50, 3, 90, 0), (3, 64, 100, 0, 75, 1),
                                               (2, 45, 33, 4, 95, 1), (2, 40, 72, 2, 100, 1)])
def test_strips_match_single_stream(oracle, world, W, H, css, q, opt):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), W, H, css, q, opt, ret), nprocs=world, join=True)
    assert ret["ok"], ret


def test_strip_rows_and_seams():
    from nvjpeg_imagecompressor_b200.strips import seam_params, strip_rows
    assert strip_rows(40000, 1, 8) == [(i * 5000, (i + 1) * 5000) for i in range(8)]
    r = strip_rows(1080, 3, 8)
    assert r[0] == (0, 144) and r[-1][1] == 1080 and all(a[1] == b[0] for a, b in zip(r, r[1:]))
    sp = seam_params([13, 8, 21], [0xAB000000, 0xCD000000, 0xEF000000])
    assert sp == [(0, 0xCD), (3, 0xEF), (3, 0xFF)]
