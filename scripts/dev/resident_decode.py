"""Three decodes of the headline JPEG with its bytes resident in HBM (b2j_decode_scan_device): run under
`ncu --metrics gpu__time_duration.sum` for the per-launch device times of that path (development aid)."""
import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import nvjpeg_imagecompressor_b200 as P
from nvjpeg_imagecompressor_b200.synth import synth
from nvjpeg_imagecompressor_b200.strips import parse_baseline_header
W, H = 8320, 40000
img = synth(W, H)
eng = P.Engine(W, H, 95, True, "422")
jpg = np.array(eng.encode(img.cpu().numpy()), copy=True)
hd = parse_baseline_header(jpg)
d_jpg = torch.from_numpy(jpg).cuda()
rec = torch.empty((H, W, 3), dtype=torch.uint8, device="cuda")
st = torch.cuda.Stream(); eng.set_stream(st.cuda_stream)
def once():
    eng.decode_scan_device(jpg[:hd["scan_off"]], d_jpg.data_ptr() + hd["scan_off"], hd["scan_end"] - hd["scan_off"], rec.data_ptr(), W * 3)
for _ in range(3):
    once()
    torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(st):
    e0.record(st)
    for _ in range(5):
        once()
    e1.record(st)
st.synchronize()
print({"lib": os.environ.get("B2J_LIB", "default"), "resident_decode_ms": round(e0.elapsed_time(e1) / 5, 3)})
