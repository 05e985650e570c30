"""Pipeline trace of gdn_ts_kernel on the g_s layer-3 shape (GPU box):  python tools/trace_gdn.py [batch]
The trace buffer is shared by every traced kernel of the launch sequence; the GDN kernel runs last, so its stamps are the ones left."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.nn as nn
from neural_image_compression_b200 import engine, _lib
from neural_image_compression_b200._lib import EPI_IGDN
from neural_image_compression_b200.gdn import GDN

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda:0")
conv = nn.ConvTranspose2d(128, 128, 5, 2, 2, output_padding=1).to(dev)
op = engine.ConvOp(conv, EPI_IGDN, gdn=GDN(128, inverse=True).to(dev))
x = engine.to_pair(torch.randn(B, 128, 192, 128, device=dev))
lib = _lib.load()
lib.nic_debug_set_trace.argtypes = [C.c_void_p]; lib.nic_debug_set_trace.restype = None
for _ in range(3):
    op.run(x, B, 128, 192, "bf16x3")
torch.cuda.synchronize()
buf = torch.zeros(148 * 32 * 16, dtype=torch.int64, device=dev)
lib.nic_debug_set_trace(buf.data_ptr())
op.run(x, B, 128, 192, "bf16x3")
torch.cuda.synchronize()
lib.nic_debug_set_trace(None)
t = buf.cpu().reshape(148, 32, 16)
names = ["w:top", "w:lo_prev_stored", "w:normalised", "-", "w:squares_next", "m:sq_seen", "-", "w:gdn_done", "w:hi_stored"]
for cta in (0, 77):
    base = int(t[cta, 0, 0])
    print(f"CTA {cta}")
    for i in range(4, 9):
        row = t[cta, i]
        print("  tile %2d: " % i + " ".join(f"{n}={int(row[j]) - base}" for j, n in enumerate(names) if n != "-" and int(row[j])))
