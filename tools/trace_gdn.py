"""Pipeline trace of gdn_x3_kernel (GPU box):  python tools/trace_gdn.py [npix]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.nn as nn
from neural_image_compression_b200 import engine, _lib
from neural_image_compression_b200._lib import EPI_GDN, LAYOUT_NCHW
from neural_image_compression_b200.gdn import GDN

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda:0")
conv = nn.Conv2d(3, 128, 5, 2, 2).to(dev)
op = engine.ConvOp(conv, EPI_GDN, gdn=GDN(128).to(dev))
x = torch.rand(B, 3, 512, 768, device=dev)
lib = _lib.load()
lib.nic_debug_set_trace.argtypes = [C.c_void_p]; lib.nic_debug_set_trace.restype = None
for _ in range(3):
    op.run(x, B, 512, 768, "bf16x3", in_layout=LAYOUT_NCHW)
torch.cuda.synchronize()
buf = torch.zeros(148 * 32 * 16, dtype=torch.int64, device=dev)
lib.nic_debug_set_trace(buf.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); op.run(x, B, 512, 768, "bf16x3", in_layout=LAYOUT_NCHW); e1.record()
torch.cuda.synchronize()
lib.nic_debug_set_trace(None)
t = buf.cpu().reshape(148, 32, 16)
print(f"first layer (conv_first_x3 + gdn_x3), B={B}: {e0.elapsed_time(e1)*1000:.1f} us")
names = ["wait_x", "x_full", "lds_done", "stage_free", "sq_done", "synced", "mma_issued", "mma_done", "out_staged", "synced2"]
for cta in (0, 77):
    base = int(t[cta, 0][t[cta, 0] > 0].min())
    print(f"CTA {cta} (clk relative to its first stamp)")
    for i in range(8):
        row = t[cta, i]
        if int(row[0]) == 0:
            break
        print("  tile %2d: " % i + "  ".join(f"{n}={int(row[j]) - base:6d}" for j, n in enumerate(names) if int(row[j]) != 0))
