import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, density_b200
from density_b200 import synth, codec
n = (int(sys.argv[1]) if len(sys.argv) > 1 else 256) << 20
C = density_b200.Chameleon
d_in = synth.synth_text(n, device="cuda")
d_out = torch.empty(C.safe_encode_buffer_size(n), dtype=torch.uint8, device="cuda")
d_sz = torch.zeros(1, dtype=torch.int64, device="cuda")
codec.encode_device("chameleon", d_in, d_out, d_sz); torch.cuda.synchronize()
m = int(d_sz.item())
d_dec = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(3):
    codec.decode_device("chameleon", d_out, m, d_dec, d_sz)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): codec.decode_device("chameleon", d_out, m, d_dec, d_sz)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"decode {n>>20} MiB: {ms:.3f} ms  {n/ms/1e6:.1f} GB/s  ok={torch.equal(d_dec, d_in)}")
