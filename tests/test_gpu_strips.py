"""The staged strip C-ABI on a real GPU: N strips of one image encoded by N contexts on ONE device (the collectives
replaced by explicit copies, as B200_PROFILING.md recommends when there are fewer GPUs than ranks); the stitched
stream must equal the single-context stream and the oracle. The torch.distributed plumbing itself is covered by
tests/test_strips_gloo.py (CPU) and exercised on real GPUs by `bench.py --gpus N`."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def encode_strips_one_gpu(img, css, q, opt, nstrips, dev_seam=False):
    from nvjpeg_imagecompressor_b200.strips import EngineBackend, seam_params, strip_rows
    H, W = img.shape[:2]
    d = torch.from_numpy(img).cuda()
    rows = strip_rows(H, css, nstrips)
    bs = [EngineBackend(W, max(b - a for a, b in rows), q, bool(opt), css, 0) for _ in rows]
    for b, (y0, y1) in zip(bs, rows):
        b.phase1(d[y0:y1].data_ptr(), W * 3, W, y1 - y0)
    torch.cuda.synchronize()
    for k, b in enumerate(bs):
        if k:
            b.pred_in.copy_(bs[k - 1].last_dc)
        else:
            b.pred_in.zero_()
    torch.cuda.synchronize()
    for b in bs:
        b.phase1b()
    torch.cuda.synchronize()
    if opt:
        tot = torch.stack([b.hist for b in bs]).sum(0).to(torch.int32)
        for b in bs:
            b.hist.copy_(tot)
    torch.cuda.synchronize()
    for b in bs:
        b.phase2(W, H)
    torch.cuda.synchronize()
    if dev_seam:   # what the NCCL path does: all-gathered (bits, first word) table stays on the device
        bits_all = torch.stack([b.strip_bits for b in bs]).contiguous()
        for k, b in enumerate(bs):
            b.phase3_dev(bits_all, k, nstrips, (1 if k == 0 else 0) | (2 if k == nstrips - 1 else 0))
    else:
        sb = np.stack([b.strip_bits.cpu().numpy() for b in bs])
        sp = seam_params(sb[:, 0], sb[:, 1].astype(np.uint64) & 0xFFFFFFFF)
        for k, b in enumerate(bs):
            b.phase3(sp[k][0], sp[k][1], (1 if k == 0 else 0) | (2 if k == nstrips - 1 else 0))
    torch.cuda.synchronize()
    parts = [b.out_view(int(b.out_len.item())).cpu().numpy() for b in bs]
    for b in bs:
        b.eng.close()
    return np.concatenate(parts)


def encode_strips_peer(img, css, q, opt, nstrips, images=3):
    """Peer-memory exchange with every 'rank' a context on this one GPU (arenas connected by raw pointers): all
    pushes are issued before any merge waits, so nothing here spins. Several images in a row exercise the sequence
    numbers and the arena sets."""
    from nvjpeg_imagecompressor_b200.strips import EngineBackend, strip_rows
    H, W = img.shape[:2]
    d = torch.from_numpy(img).cuda()
    rows = strip_rows(H, css, nstrips)
    bs = [EngineBackend(W, max(b - a for a, b in rows), q, bool(opt), css, 0) for _ in rows]
    arenas = [b.eng.peer_export()[1] for b in bs]
    for k, b in enumerate(bs):
        b.eng.peer_connect(k, nstrips, arenas)
    outs = []
    for it in range(images):
        for b, (y0, y1) in zip(bs, rows):
            b.phase1x(d[y0:y1].data_ptr(), W * 3, W, y1 - y0)
        torch.cuda.synchronize()
        for k, b in enumerate(bs):
            b.phase2x(None, k, nstrips, W, H, (1 if k == 0 else 0) | (2 if k == nstrips - 1 else 0))
        torch.cuda.synchronize()
        outs.append(np.concatenate([b.out_view(b.eng.encode_finish()).cpu().numpy() for b in bs]))
    for b in bs:
        b.eng.close()
    return outs


def encode_strips_one_collective(img, css, q, opt, nstrips):
    """phase1x -> (all_gather of the records) -> phase2x: the schedule StripEncoder runs over NCCL."""
    from nvjpeg_imagecompressor_b200.strips import EngineBackend, strip_rows
    H, W = img.shape[:2]
    d = torch.from_numpy(img).cuda()
    rows = strip_rows(H, css, nstrips)
    bs = [EngineBackend(W, max(b - a for a, b in rows), q, bool(opt), css, 0) for _ in rows]
    for b, (y0, y1) in zip(bs, rows):
        b.phase1x(d[y0:y1].data_ptr(), W * 3, W, y1 - y0)
    torch.cuda.synchronize()
    rec_all = torch.stack([b.record for b in bs]).contiguous()
    torch.cuda.synchronize()
    for k, b in enumerate(bs):
        b.phase2x(rec_all, k, nstrips, W, H, (1 if k == 0 else 0) | (2 if k == nstrips - 1 else 0))
    torch.cuda.synchronize()
    parts = []
    for b in bs:
        n = b.eng.encode_finish()          # raises when a device check (predicted vs coded bit count) failed
        parts.append(b.out_view(n).cpu().numpy())
        b.eng.close()
    return np.concatenate(parts)


@pytest.mark.parametrize("W,H,css,q,opt,n", [(256, 320, 1, 95, 1, 2), (256, 320, 1, 95, 1, 8), (200, 333, 3, 90, 1, 4),
                                           (129, 200, 0, 75, 0, 3), (96, 250, 2, 95, 1, 5), (160, 64, 4, 100, 1, 8)])
def test_strips_equal_single_stream(oracle, W, H, css, q, opt, n):
    img = oracle.synth(W, H, 7, 8)
    want = oracle.encode(img, css, q, opt)
    for dev_seam in (False, True):
        out = encode_strips_one_gpu(img, css, q, opt, n, dev_seam)
        assert out.size == want.size and np.array_equal(out, want), f"dev_seam={dev_seam}"
    out = encode_strips_one_collective(img, css, q, opt, n)
    assert out.size == want.size and np.array_equal(out, want), "one-collective schedule"
    for i, out in enumerate(encode_strips_peer(img, css, q, opt, n, images=6)):
        assert out.size == want.size and np.array_equal(out, want), f"peer-memory exchange, image {i}"


def test_peer_exchange_times_out_instead_of_hanging():
    """A rank whose peers never push gets error 9 from the bounded wait, not a hung GPU."""
    import os
    import time
    os.environ["B2J_PEER_TIMEOUT_MS"] = "1500"   # read by b2j_peer_connect
    import nvjpeg_imagecompressor_b200 as P
    from nvjpeg_imagecompressor_b200.strips import EngineBackend
    img = torch.zeros((64, 64, 3), dtype=torch.uint8, device="cuda")
    b0, b1 = (EngineBackend(64, 32, 90, True, 1, 0) for _ in range(2))
    arenas = [b.eng.peer_export()[1] for b in (b0, b1)]
    b0.eng.peer_connect(0, 2, arenas)
    b1.eng.peer_connect(1, 2, arenas)
    b0.phase1x(img.data_ptr(), 64 * 3, 64, 32)      # rank 1 never encodes
    t0 = time.time()
    b0.phase2x(None, 0, 2, 64, 64, 1)
    with pytest.raises(P.B2JError):
        b0.eng.encode_finish()
    assert 1.0 < time.time() - t0 < 20
    del os.environ["B2J_PEER_TIMEOUT_MS"]
    b0.eng.close(); b1.eng.close()


def test_strips_headline_slab(oracle, golden):
    import hashlib
    img = oracle.synth(8320, 2000, 0, 8)
    c = golden["slab"][0]
    out = encode_strips_one_gpu(img, c["css"], c["quality"], c["optimize"], 8, dev_seam=True)
    assert out.size == c["jpeg_len"] and hashlib.sha256(out.tobytes()).hexdigest() == c["jpeg_sha256"]
    out = encode_strips_one_collective(img, c["css"], c["quality"], c["optimize"], 8)
    assert out.size == c["jpeg_len"] and hashlib.sha256(out.tobytes()).hexdigest() == c["jpeg_sha256"]


def test_one_collective_flat_and_tiny_strips(oracle):
    """Constant image (every code 1-2 bits, seams inside the first tokens) and strips of a single MCU row."""
    for img, css, n in ((np.full((64, 48, 3), 77, np.uint8), 1, 4), (oracle.synth(40, 64, 3, 8), 0, 8),
                        (np.zeros((128, 16, 3), np.uint8), 3, 8)):
        for opt in (0, 1):
            want = oracle.encode(img, css, 90, opt)
            out = encode_strips_one_collective(img, css, 90, opt, n)
            assert out.size == want.size and np.array_equal(out, want), (img.shape, css, opt)


@pytest.mark.parametrize("css", ["444", "422", "411"])
def test_strip_decodes_alone_to_the_images_rows(oracle, css):
    """No vertical subsampling: a strip encoded and decoded as an image of its own = its rows of the whole decode."""
    import nvjpeg_imagecompressor_b200 as P
    img = oracle.synth(200, 160, 5, 8)
    full = P.Engine(200, 160, 90, True, css)
    whole = full.decode(full.encode(img))
    for y0, y1 in ((0, 40), (40, 104), (104, 160)):
        e = P.Engine(200, y1 - y0, 90, True, css)
        part = e.decode(e.encode(np.ascontiguousarray(img[y0:y1])))
        assert np.array_equal(part, whole[y0:y1]), (css, y0, y1)
        e.close()
    full.close()


@pytest.mark.parametrize("css,W,H,n", [("420", 200, 160, 3), ("440", 131, 203, 4), ("420", 64, 250, 8), ("422", 96, 64, 2)])
def test_strips_reconstruct_with_chroma_halos(oracle, css, W, H, n):
    """Vertical subsampling: the strips' planes + one chroma row swapped across every border (what StripSecondary sends
    between ranks) + the colour step = the whole image's decode, bit for bit."""
    import nvjpeg_imagecompressor_b200 as P
    from nvjpeg_imagecompressor_b200 import _native as N
    from nvjpeg_imagecompressor_b200.strips import _view, strip_rows
    img = oracle.synth(W, H, 11, 8)
    whole = oracle.decode(oracle.encode(img, N.CSS[css], 90, 1))
    d = torch.from_numpy(img).cuda()
    rows = [r for r in strip_rows(H, N.CSS[css], n) if r[1] > r[0]]
    engs, rps = [], []
    for y0, y1 in rows:
        e = P.Engine(W, y1 - y0, 90, True, css)
        e.set_debug(1)
        e.encode_device(d[y0:y1].data_ptr(), W * 3, W, y1 - y0)
        engs.append(e)
        rps.append(e.reconstruct_planes())
    torch.cuda.synchronize()
    v = lambda p, rp: _view(p, (int(rp.row_bytes),), "|u1", d.device)
    for k in range(len(rows) - 1):
        a, b = rps[k], rps[k + 1]
        v(a.cb_halo_bottom, a).copy_(v(b.cb_first, b)); v(a.cr_halo_bottom, a).copy_(v(b.cr_first, b))
        v(b.cb_halo_top, b).copy_(v(a.cb_last, a)); v(b.cr_halo_top, b).copy_(v(a.cr_last, a))
    torch.cuda.synchronize()
    out = torch.empty((H, W, 3), dtype=torch.uint8, device=d.device)
    for k, (e, (y0, y1)) in enumerate(zip(engs, rows)):
        e.reconstruct_color(out[y0:y1].data_ptr(), W * 3, k > 0, k + 1 < len(rows))
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), whole)
    for e in engs:
        e.close()


def test_strip_secondary_world1_equals_b2j_secondary(oracle):
    import nvjpeg_imagecompressor_b200 as P
    from nvjpeg_imagecompressor_b200.strips import StripSecondary
    W, H = 256, 192
    img = oracle.synth(W, H, 9, 8)
    eng = P.Engine(W, H, 95, True, "422")
    j1, j2, recon, psnr = eng.secondary(img, diff_mode=1)
    sec = StripSecondary(W, H, 95, True, "422", diff_mode=1, rank=0, world=1, device=0)
    d = torch.from_numpy(img).cuda()
    n1, n2, ps = sec.run(d.data_ptr(), W * 3)
    torch.cuda.synchronize()
    a = sec.enc1.gather_jpeg(0).cpu().numpy()
    b = sec.enc2.gather_jpeg(0).cpu().numpy()
    assert np.array_equal(a, j1) and np.array_equal(b, j2)
    assert abs(ps - psnr) < 1e-9
    assert np.array_equal(sec.recon[:H].cpu().numpy(), recon)
    eng.close()


@pytest.mark.parametrize("W,H,css,q,opt,n", [(256, 320, 1, 95, 1, 2), (200, 333, 3, 90, 1, 4), (129, 200, 0, 75, 0, 3),
                                           (1040, 2600, 1, 95, 1, 8), (64, 16, 1, 95, 1, 4)])
def test_multi_context_one_process(oracle, W, H, css, q, opt, n):
    """b2j_multi_*: several contexts driven by ONE process and one host thread (here all on this GPU; on a multi-GPU
    box one per GPU), host image in -> host JPEG out, the records exchanged through peer memory: the bytes are the
    single-stream bytes, with pageable and with pinned host buffers, several images in a row."""
    import nvjpeg_imagecompressor_b200 as P
    img = oracle.synth(W, H, 4, 8)
    want = oracle.encode(img, css, q, opt)
    m = P.MultiEngine(W, H, q, bool(opt), css, devices=[0] * n)
    for it in range(3):
        got = m.encode(img)                                   # pageable numpy memory
        assert got.size == want.size and np.array_equal(got, want), (it, "pageable")
    h_img = torch.from_numpy(img).pin_memory()
    h_out = torch.empty(W * H * 3 + 65536, dtype=torch.uint8).pin_memory()
    nb = m.encode_ptr(h_img.data_ptr(), W * 3, W, H, h_out.data_ptr(), h_out.numel())
    assert nb == want.size and np.array_equal(h_out[:nb].numpy(), want), "pinned"
    m.close()


@pytest.mark.parametrize("W,H,css,q,rows,n", [(640, 720, 1, 95, 1, 2), (640, 720, 3, 90, 1, 4), (333, 500, 2, 95, 2, 3),
                                            (1040, 2000, 1, 100, 1, 8), (200, 900, 0, 75, 3, 5), (64, 64, 3, 95, 1, 4)])
def test_single_image_decode_over_strips(oracle, W, H, css, q, rows, n):
    """one image with restart intervals of whole MCU rows decoded by N ranks (here N decoders in one
    process; scripts/decode_multi.py runs them as N processes on N GPUs): every rank takes a byte range of the scan,
    finds its RSTn markers on the GPU, the ranks share only their marker counts, each decodes its intervals plus a
    halo interval; the rows put together are libjpeg-turbo's pixels (4:2:0 / 4:4:0 included: vertical chroma filter)."""
    import nvjpeg_imagecompressor_b200 as P
    from nvjpeg_imagecompressor_b200.strips import StripDecoder
    img = oracle.synth(W, H, 9, 8)
    g = oracle.geometry(W, H, css)
    jpg = oracle.encode(img, css, q, 1, rows * g.mcux)
    want = oracle.decode(jpg)
    decs = [StripDecoder(jpg, r, n) for r in range(n)]
    counts = [d.count_markers() for d in decs]           # what one all_gather of an int per rank gives every rank
    assert sum(counts) == -(-g.mcuy // rows) - 1
    out = np.zeros_like(want)
    covered = 0
    for d in decs:
        y0, t = d.decode(counts)
        out[y0:y0 + t.shape[0]] = t.cpu().numpy()
        covered += t.shape[0]
        d.close()
    assert covered == H
    assert np.array_equal(out, want)
