"""Diagnose the copy-map iteration of the Chameleon encoder on text with incompressible bursts."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, density_b200
from density_b200 import synth, codec
C = density_b200.Chameleon
def run(n_mib, offs, same_seed, burst=65536):
    n = n_mib << 20
    d_in = synth.synth_text(n, device="cuda")
    for k, off in enumerate(offs):
        d_in[off:off + burst] = synth.random_bytes(burst, 99 if same_seed else 1000 + k, device="cuda")
    d_enc = torch.empty(C.safe_encode_buffer_size(n), dtype=torch.uint8, device="cuda")
    d_sz = torch.zeros(1, dtype=torch.int64, device="cuda")
    codec.encode_device("chameleon", d_in, d_enc, d_sz); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); codec.encode_device("chameleon", d_in, d_enc, d_sz); e1.record(); torch.cuda.synchronize()
    st = (ctypes.c_uint64 * 6)(); density_b200.load().density_b200_encode_status(st)
    print(f"n={n_mib} MiB bursts={len(offs)} x {burst} same_seed={same_seed}: {e0.elapsed_time(e1):.2f} ms  status out,nonquiet,err,first_nq,converged,chg = {list(st)}", flush=True)
M = 1 << 20
run(128, [8 * M], True)
run(128, [8 * M], True, burst=4096)
run(128, [8 * M, 72 * M], True)
run(128, [8 * M, 72 * M], False)
run(128, [8 * M + 256 * 7], True)
run(512, [8 * M + k * 64 * M for k in range(8)], False)
# the pathological chain: 16 identical bursts in 1 GiB; reference-facing (blocking) entry point vs stream-ordered auto path
import time
n = 1024 << 20
d_in = synth.synth_text(n, device="cuda")
for k in range(16):
    d_in[(8 + 64 * k) * M:(8 + 64 * k) * M + 65536] = synth.random_bytes(65536, 99, device="cuda")
d_enc = torch.empty(C.safe_encode_buffer_size(n), dtype=torch.uint8, device="cuda")
C.encode(d_in, d_enc); torch.cuda.synchronize()
t = time.perf_counter(); m = C.encode(d_in, d_enc); torch.cuda.synchronize(); dt = (time.perf_counter() - t) * 1e3
st = (ctypes.c_uint64 * 6)(); density_b200.load().density_b200_encode_status(st)
print(f"16 identical bursts, 1 GiB, chameleon_encode() on device pointers (blocking iteration): {dt:.1f} ms  {n/dt/1e6:.2f} GB/s  status {list(st)}")
