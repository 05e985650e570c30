"""Pipeline trace of first_fused_x3_kernel (GPU box):  python tools/trace_first.py"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.nn as nn
from neural_image_compression_b200 import engine, _lib
from neural_image_compression_b200._lib import EPI_GDN, LAYOUT_NCHW
from neural_image_compression_b200.gdn import GDN

B = 16
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
op.run(x, B, 512, 768, "bf16x3", in_layout=LAYOUT_NCHW)
torch.cuda.synchronize()
lib.nic_debug_set_trace(None)
t = buf.cpu().reshape(148, 32, 16)
names = ["w:top", "-", "w:normalised", "-", "w:squares_next", "-", "-", "w:gdn_done", "w:hi_stored", "-", "w:lo_prev_stored",
         "m:conv_next", "m:sq_seen", "p:patch_ready", "p:a_full", "p:a_empty_seen"]
for cta in (0, 77):
    base = int(t[cta, 0, 0])
    print(f"CTA {cta}")
    for i in range(4, 9):
        row = t[cta, i]
        print("  tile %2d: " % i + " ".join(f"{n}={int(row[j]) - base}" for j, n in enumerate(names) if int(row[j])))
