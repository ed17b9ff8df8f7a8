import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, density_b200, numpy as np
from density_b200 import synth, codec
n = (int(sys.argv[1]) if len(sys.argv) > 1 else 256) << 20
kinds = sys.argv[2].split(",") if len(sys.argv) > 2 else ["mixed", "random"]
C = density_b200.Chameleon
for kind in kinds:
    d_in = synth.synth_mixed(n, device="cuda") if kind == "mixed" else synth.random_bytes(n, 5, device="cuda")
    d_out = torch.empty(C.safe_encode_buffer_size(n), dtype=torch.uint8, device="cuda")
    d_sz = torch.zeros(1, dtype=torch.int64, device="cuda")
    codec.encode_device("chameleon", d_in, d_out, d_sz); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): codec.encode_device("chameleon", d_in, d_out, d_sz)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    m = int(d_sz.item())
    import ctypes
    st = np.zeros(8, dtype=np.uint32)
    print(f"{kind} encode {n>>20} MiB: {ms:.3f} ms  {n/ms/1e6:.2f} GB/s ratio {n/m:.3f} fast={density_b200.load().density_b200_last_encode_was_fast()}")
    if "--check" in sys.argv:
        import oracle
        want = oracle.encode("chameleon", d_in.cpu().numpy())
        got = d_out[:m].cpu().numpy()
        print("  bit-exact vs oracle:", got.size == want.size and bool((got == want).all()))
