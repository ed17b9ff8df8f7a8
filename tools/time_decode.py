"""Chameleon decode, decode pass kernel 1 vs 7: CUDA-event time of the whole decode (path 1 = parallel only) and round-trip check."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, density_b200
from density_b200 import synth, codec
n = (int(sys.argv[1]) if len(sys.argv) > 1 else 1024) << 20
kinds = sys.argv[2].split(",") if len(sys.argv) > 2 else ["text"]
impls = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1, 7]
lib = density_b200.load()
C = density_b200.Chameleon
for kind in kinds:
    d_in = {"text": synth.synth_text, "mixed": synth.synth_mixed}.get(kind, lambda n, device: synth.random_bytes(n, 5, device=device))(n, device="cuda")
    d_enc = torch.empty(C.safe_encode_buffer_size(n), dtype=torch.uint8, device="cuda")
    d_sz = torch.zeros(1, dtype=torch.int64, device="cuda")
    codec.encode_device("chameleon", d_in, d_enc, d_sz); torch.cuda.synchronize()
    m = int(d_sz.item())
    d_dec = torch.empty(n, dtype=torch.uint8, device="cuda")
    for impl in impls:
        lib.density_b200_test_set_decode_impl(impl)
        d_dec.zero_()
        codec.decode_device("chameleon", d_enc, m, d_dec, d_sz, path=1); torch.cuda.synchronize()
        ok = int(d_sz.item()) == n and bool(torch.equal(d_dec, d_in))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): codec.decode_device("chameleon", d_enc, m, d_dec, d_sz, path=1)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"{kind} {n >> 20} MiB (stream {m >> 20} MiB) decode impl {impl}: {ms:.3f} ms  {n / ms / 1e6:.1f} GB/s  round trip {'OK' if ok else 'MISMATCH'}", flush=True)
        if not ok:
            neq = (d_dec != d_in).nonzero()
            print("   d_sz", int(d_sz.item()), "mismatching bytes", int(neq.numel()), "first", int(neq[0]) if neq.numel() else None)
lib.density_b200_test_set_decode_impl(7)
