"""Chameleon decode of a mostly-text stream with a few copy-mode episodes (random bursts every 64 MiB): boundary walk with chunk jumps."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, density_b200
from density_b200 import synth, codec
n = next((int(a) for a in sys.argv[1:] if a.isdigit()), 1024) << 20
C = density_b200.Chameleon
d_in = synth.synth_text(n, device="cuda")
for off in range(8 << 20, n, 64 << 20):
    d_in[off:off + 65536] = synth.random_bytes(65536, 99, device="cuda")
d_enc = torch.empty(C.safe_encode_buffer_size(n), dtype=torch.uint8, device="cuda")
d_sz = torch.zeros(1, dtype=torch.int64, device="cuda")
for _ in range(2): codec.encode_device("chameleon", d_in, d_enc, d_sz)
torch.cuda.synchronize()
m = int(d_sz.item())
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(3): codec.encode_device("chameleon", d_in, d_enc, d_sz)
t1.record(); torch.cuda.synchronize()
ems = t0.elapsed_time(t1) / 3
print(f"chameleon encode, text with bursts, {n>>20} MiB: {ems:.3f} ms  {n/ems/1e6:.2f} GB/s (copy-map fixed-point iteration active)")
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
print(f"chameleon decode, text with {len(range(8 << 20, n, 64 << 20))} random bursts, {n>>20} MiB: {ms:.3f} ms  {n/ms/1e6:.2f} GB/s  in-order boundaries={stat[6]}  round trip {'OK' if ok else 'MISMATCH'}")
