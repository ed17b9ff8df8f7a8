"""Device-resident throughput of every codec entry point (dev tool; the contract benchmark is bench.py)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, density_b200
from density_b200 import synth, codec

def run(alg, nbytes, kind="text", reps=3):
    C = density_b200.CODECS[alg]
    d_in = synth.synth_text(nbytes, device="cuda") if kind == "text" else synth.synth_mixed(nbytes, device="cuda")
    d_out = torch.empty(C.safe_encode_buffer_size(nbytes), dtype=torch.uint8, device="cuda")
    d_sz = torch.zeros(1, dtype=torch.int64, device="cuda")
    d_dec = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    d_sz2 = torch.zeros(1, dtype=torch.int64, device="cuda")
    def t(fn):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    ms_e = t(lambda: codec.encode_device(alg, d_in, d_out, d_sz))
    m = int(d_sz.item())
    ms_d = t(lambda: codec.decode_device(alg, d_out, m, d_dec, d_sz2))
    ok = int(d_sz2.item()) == nbytes and torch.equal(d_dec, d_in)
    print(f"{alg:9s} {kind:5s} {nbytes>>20:5d} MiB  ratio {nbytes/m:5.3f}  encode {ms_e:9.3f} ms {nbytes/ms_e/1e6:9.2f} GB/s   decode {ms_d:9.3f} ms {nbytes/ms_d/1e6:9.2f} GB/s  roundtrip {'ok' if ok else 'FAIL'}")

if __name__ == "__main__":
    size = int(sys.argv[1]) << 20 if len(sys.argv) > 1 else 16 << 20
    for alg in ("chameleon", "cheetah", "lion"):
        run(alg, size, "text")
    run("chameleon", size, "mixed")
