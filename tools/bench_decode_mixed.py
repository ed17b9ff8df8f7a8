"""Chameleon decode of a stream with copy-mode blocks (synthetic mixed text/binary), device-resident: boundary walk + dictionary passes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, density_b200
from density_b200 import synth, codec
n = next((int(a) for a in sys.argv[1:] if a.isdigit()), 256) << 20
C = density_b200.Chameleon
d_in = synth.synth_mixed(n, device="cuda")
d_enc = torch.empty(C.safe_encode_buffer_size(n), dtype=torch.uint8, device="cuda")
d_sz = torch.zeros(1, dtype=torch.int64, device="cuda")
codec.encode_device("chameleon", d_in, d_enc, d_sz)
torch.cuda.synchronize()
m = int(d_sz.item())
d_dec = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(2): codec.decode_device("chameleon", d_enc, m, d_dec, d_sz, path=1)
torch.cuda.synchronize()
ok = int(d_sz.item()) == n and torch.equal(d_dec, d_in)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): codec.decode_device("chameleon", d_enc, m, d_dec, d_sz, path=1)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
import ctypes
stat = (ctypes.c_uint64 * 10)()
density_b200.load().density_b200_decode_status(stat)
print("  status: out_bytes main_blocks tail_off nonquiet error last_inc seq penalty start prev =", list(stat))
if not ok:
    neq = (d_dec != d_in).nonzero()
    print("  d_sz", int(d_sz.item()), "n", n, "mismatching bytes", int(neq.numel()), "first", int(neq[0]) if neq.numel() else None, "last", int(neq[-1]) if neq.numel() else None)
print(f"chameleon decode, mixed {n>>20} MiB (stream {m>>20} MiB): {ms:.3f} ms  {n/ms/1e6:.2f} GB/s  round trip {'OK' if ok else 'MISMATCH'}")
