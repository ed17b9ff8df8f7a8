"""Chameleon encode, flag pass kernel 1 vs 6: CUDA-event time of the whole encode, output equality, oracle check (optional)."""
import sys, os, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, density_b200
from density_b200 import synth, codec
n = (int(sys.argv[1]) if len(sys.argv) > 1 else 1024) << 20
kinds = sys.argv[2].split(",") if len(sys.argv) > 2 else ["text"]
impls = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1, 6]
lib = density_b200.load()
C = density_b200.Chameleon
for kind in kinds:
    d_in = {"text": synth.synth_text, "mixed": synth.synth_mixed}.get(kind, lambda n, device: synth.random_bytes(n, 5, device=device))(n, device="cuda")
    d_out = torch.empty(C.safe_encode_buffer_size(n), dtype=torch.uint8, device="cuda")
    d_sz = torch.zeros(1, dtype=torch.int64, device="cuda")
    digests = {}
    for impl in impls:
        lib.density_b200_test_set_flag_impl(impl)
        d_out.zero_()
        codec.encode_device("chameleon", d_in, d_out, d_sz); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        lib.density_b200_profile_enable(1)
        e0.record()
        for _ in range(5): codec.encode_device("chameleon", d_in, d_out, d_sz)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        import ctypes
        prof = (ctypes.c_float * 3)(); lib.density_b200_profile_get(prof); lib.density_b200_profile_enable(0)
        m = int(d_sz.item())
        digests[impl] = (m, hashlib.sha256(d_out[:m].cpu().numpy().tobytes()).hexdigest())
        print(f"{kind} {n >> 20} MiB impl {impl}: {ms:.3f} ms  {n / ms / 1e6:.1f} GB/s  out {m}  flag {prof[0]:.3f} mid {prof[1]:.3f} emit {prof[2]:.3f} ms  fast={lib.density_b200_last_encode_was_fast()}", flush=True)
    print("  impls agree:", len(set(digests.values())) == 1)
    if "--check" in sys.argv:
        import oracle
        want = oracle.encode("chameleon", d_in.cpu().numpy())
        print("  oracle:", (want.size, hashlib.sha256(want.tobytes()).hexdigest()) == digests[impls[-1]])
lib.density_b200_test_set_flag_impl(6)
