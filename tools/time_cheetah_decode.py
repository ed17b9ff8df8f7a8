"""One Cheetah encode + parallel decode of the 1 GiB bench text (for `ncu --metrics gpu__time_duration.sum` launch lists and round counts)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import density_b200
from density_b200 import synth
n = int(sys.argv[1]) << 20 if len(sys.argv) > 1 else 1 << 30
kind = sys.argv[2] if len(sys.argv) > 2 else "text"
d_in = synth.synth_text(n, device="cuda") if kind == "text" else synth.synth_mixed(n, device="cuda")
C = density_b200.Cheetah
d_enc = torch.empty(C.safe_encode_buffer_size(n), dtype=torch.uint8, device="cuda")
d_sz = torch.zeros(1, dtype=torch.int64, device="cuda")
density_b200.encode_device("cheetah", d_in, d_enc, d_sz, path=0)
torch.cuda.synchronize()
m = int(d_sz.item())
d_dec = torch.empty(n, dtype=torch.uint8, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(2):
    e0.record()
    density_b200.decode_device("cheetah", d_enc, m, d_dec, d_sz, path=1)
    e1.record(); torch.cuda.synchronize()
    r = (ctypes.c_uint32 * 4)()
    density_b200.load().density_b200_cheetah_decode_rounds(r)
    print(f"{kind} {n >> 20} MiB: stream {m} B, decode {e0.elapsed_time(e1):.2f} ms = {n / e0.elapsed_time(e1) / 1e6:.2f} GB/s, rounds {r[0]}, settled {r[1]}, ok {int(d_sz.item()) == n and bool(torch.equal(d_dec, d_in))}")
