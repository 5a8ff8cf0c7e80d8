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


TOK_RAWDC = 1 << 26


def _nbits(v):
    return int(abs(int(v))).bit_length()


def _vbits(v, nb):
    v = int(v)
    return (v if v >= 0 else v - 1) & ((1 << nb) - 1)


class OracleBackendX(OracleBackend):
    """The one-exchange schedule (include/b2jpeg.h b2j_strip_record, csrc/enc_huff.cu k_strip_merge / k_strip_seam)
    restated with numpy: own symbol counts without the first MCU's DC symbols, first/last DCs and the first 8 tokens
    go out; every strip's bit count comes back as counts x (code length + value bits)."""

    def __init__(self, O, quality, optimize, css):
        super().__init__(O, quality, optimize, css)
        from nvjpeg_imagecompressor_b200 import _native as N
        self.record = torch.zeros(N.STRIP_RECORD_BYTES, dtype=torch.uint8)

    # record fields as numpy views of a (world, bytes) uint8 array
    @staticmethod
    def _fields(rec):
        rec = np.ascontiguousarray(rec)
        hist = rec[..., :4112].view(np.uint32).reshape(rec.shape[:-1] + (4, 257))
        first = rec[..., 4112:4120].view(np.int16)
        last = rec[..., 4120:4128].view(np.int16)
        tok = rec[..., 4128:4160].view(np.uint32)
        ntok = rec[..., 4160:4164].view(np.uint32)
        return hist, first, last, tok, ntok

    def _first_tokens(self, limit=8):
        """Engine token format: [29:28] ZRL count, [25:24] table, [23:16] symbol, [15:0] value bits; the first block of a
        component in the strip carries its raw DC (TOK_RAWDC | table << 24 | component << 16 | dc)."""
        b, toks, prev = self.g.bpm, [], {}
        blocks = self.coef.reshape(-1, 64)
        for i in range(min(len(blocks), 4 * b)):
            k = i % b
            comp = 0 if k < b - 2 else k - (b - 3)
            tdc, tac = (0, 1) if comp == 0 else (2, 3)
            dc = int(blocks[i, 0])
            if comp not in prev:
                toks.append(TOK_RAWDC | (tdc << 24) | (comp << 16) | (dc & 0xFFFF))
            else:
                d = dc - prev[comp]; nb = _nbits(d)
                toks.append((tdc << 24) | (nb << 16) | _vbits(d, nb))
            prev[comp] = dc
            run = 0
            for z in range(1, 64):
                v = int(blocks[i, z])
                if v == 0:
                    run += 1
                    continue
                nb = _nbits(v)
                toks.append(((run >> 4) << 28) | (tac << 24) | ((((run & 15) << 4) | nb) << 16) | _vbits(v, nb))
                run = 0
            if run:
                toks.append(tac << 24)   # EOB
            if len(toks) >= limit:
                break
        return toks[:limit]

    def phase1x(self, img, step, W, rows):
        self.phase1(img, step, W, rows)
        O, b = self.O, self.g.bpm
        h = O.histogram(self.coef, b, np.zeros(3, np.int16)).astype(np.int64)
        blocks = self.coef.reshape(-1, 64)
        firsts = [int(blocks[0, 0]), int(blocks[b - 2, 0]), int(blocks[b - 1, 0])]
        for c, dc in enumerate(firsts):           # the first MCU's DC symbols wait for the previous strip's DCs
            h[0 if c == 0 else 2][_nbits(dc)] -= 1
        rec = np.zeros(self.record.numel(), np.uint8)
        hist, first, last, tok, ntok = self._fields(rec)
        hist[...] = h.astype(np.uint32)
        first[:3] = firsts
        last[:3] = self.last_dc[:3].numpy()
        t = self._first_tokens()
        tok[: len(t)] = t
        ntok[0] = len(t)
        self.record.copy_(torch.from_numpy(rec))

    def phase2x(self, records_all, rank, world, W, H, flags):
        O = self.O
        hist, first, last, tok, ntok = self._fields(records_all.numpy())
        pred = lambda k: last[k - 1][:3].astype(np.int64) if k else np.zeros(3, np.int64)
        total = hist.astype(np.int64).sum(0)
        for k in range(world):                     # DC symbols at the strip starts
            for c in range(3):
                total[0 if c == 0 else 2][_nbits(int(first[k][c]) - int(pred(k)[c]))] += 1
        self.pred_in[:3] = torch.from_numpy(pred(rank).astype(np.int16))
        self.hist.copy_(torch.from_numpy(total.astype(np.int32).ravel()))
        self.phase2(W, H)                          # tables from the merged counts + this strip's bits at phase 0
        size = [O.derive_codes(self.bits[t], self.vals[t])[1].astype(np.int64) for t in range(4)]
        code = [O.derive_codes(self.bits[t], self.vals[t])[0].astype(np.int64) for t in range(4)]
        extra = [np.arange(256), np.arange(256) & 15, np.arange(256), np.arange(256) & 15]

        def strip_bits(k):
            n = sum(int((hist[k][t][:256].astype(np.int64) * (size[t] + extra[t])).sum()) for t in range(4))
            for c in range(3):
                nb = _nbits(int(first[k][c]) - int(pred(k)[c]))
                n += int(size[0 if c == 0 else 2][nb]) + nb
            return n
        assert strip_bits(rank) == int(self.strip_bits[0]), "predicted bit count != coded bit count"
        G = sum(strip_bits(k) for k in range(rank))
        skip, ext = (8 - G % 8) % 8, 0xFF
        if rank + 1 < world:                       # head of the next strip's bit string from its first tokens
            acc = n = 0
            for tk in tok[rank + 1][: int(ntok[rank + 1][0])]:
                tk = int(tk)
                if tk & TOK_RAWDC:
                    c = (tk >> 16) & 3
                    dc = ((tk & 0xFFFF) ^ 0x8000) - 0x8000
                    d = dc - int(last[rank][c]); nb = _nbits(d)
                    tk = ((tk >> 24) & 3) << 24 | (nb << 16) | _vbits(d, nb)
                t, sym = (tk >> 24) & 3, (tk >> 16) & 0xFF
                for _ in range((tk >> 28) & 3):
                    acc = (acc << int(size[t][0xF0])) | int(code[t][0xF0]); n += int(size[t][0xF0])
                nv = sym & 15 if t & 1 else sym
                acc = (acc << int(size[t][sym])) | int(code[t][sym]); n += int(size[t][sym])
                acc = (acc << nv) | (tk & 0xFFFF); n += nv
                if n >= 8:
                    break
            assert n >= 8
            ext = (acc >> (n - 8)) & 0xFF
        self.phase3(skip, ext, flags)


def _worker(rank, world, port, W, H, css, q, opt, ret, one_exchange=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle as O
    from nvjpeg_imagecompressor_b200.strips import StripEncoder
    img = O.synth(W, H, 3, 8)
    enc = StripEncoder(W, H, q, opt, css, backend=(OracleBackendX if one_exchange else OracleBackend)(O, q, opt, css))
    if one_exchange:   # the CPU backend cannot map peer memory: every rank falls back to the all_gather of the records
        assert enc.exchange == "one all_gather", enc.exchange
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


@pytest.mark.parametrize("world,W,H,css,q,opt", [(2, 96, 80, 1, 95, 1), (3, 64, 100, 0, 75, 1), (2, 70, 50, 3, 90, 0),
                                               (3, 45, 33, 4, 95, 1)])
def test_one_exchange_schedule_matches_single_stream(oracle, world, W, H, css, q, opt):
    """phase1x -> ONE all_gather of the strips' records -> phase2x, restated on the CPU (OracleBackendX)."""
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), W, H, css, q, opt, ret, True), nprocs=world, join=True)
    assert ret["ok"], ret


def test_strip_rows_and_seams():
    from nvjpeg_imagecompressor_b200.strips import seam_params, strip_rows
    assert strip_rows(40000, 1, 8) == [(i * 5000, (i + 1) * 5000) for i in range(8)]
    r = strip_rows(1080, 3, 8)
    assert r[0] == (0, 144) and r[-1][1] == 1080 and all(a[1] == b[0] for a, b in zip(r, r[1:]))
    sp = seam_params([13, 8, 21], [0xAB000000, 0xCD000000, 0xEF000000])
    assert sp == [(0, 0xCD), (3, 0xEF), (3, 0xFF)]
