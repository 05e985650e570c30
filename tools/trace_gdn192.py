"""Pipeline trace of gdn_x3c_kernel<3> (GDN at 192 channels, gamma streamed) (GPU box):  python tools/trace_gdn192.py"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.nn as nn
from neural_image_compression_b200 import engine, _lib
from neural_image_compression_b200._lib import EPI_GDN
from neural_image_compression_b200.gdn import GDN

dev = torch.device("cuda:0")
conv = nn.Conv2d(192, 192, 5, 2, 2).to(dev)
op = engine.ConvOp(conv, EPI_GDN, gdn=GDN(192).to(dev))
B, h, w = 4, 384, 512
x = engine.to_pair(torch.randn(B, h, w, 192, device=dev))
lib = _lib.load()
lib.nic_debug_set_trace.argtypes = [C.c_void_p]; lib.nic_debug_set_trace.restype = None
for _ in range(2):
    op.run(x, B, h, w, "bf16x3")
torch.cuda.synchronize()
buf = torch.zeros(148 * 32 * 16, dtype=torch.int64, device=dev)
lib.nic_debug_set_trace(buf.data_ptr())
op.run(x, B, h, w, "bf16x3")
torch.cuda.synchronize()
lib.nic_debug_set_trace(None)
t = buf.cpu().reshape(148, 32, 16)
names = ["top", "x_full", "squared", "stores_read", None, "mma_done", "applied"]
for cta in (0, 77):
    print(f"CTA {cta} (the conv kernel writes the same buffer first: only the GDN kernel's stamps survive)")
    base = int(t[cta, 2, 0])
    for i in range(2, 7):
        row = t[cta, i]
        print("  tile %2d: " % i + " ".join(f"{n}={int(row[j]) - base}" for j, n in enumerate(names) if n))
