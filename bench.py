#!/usr/bin/env python
"""bench.py — headline benchmark of the density_b200 hot path (see DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU (oracle port)

A "step" is one Chameleon encode of one synthetic-text buffer (BASELINE.json configs[1]: 1 GiB per GPU, resident in HBM
when the timed region starts). Metric: input GB/s (uncompressed bytes / time, GB = 1e9 B — the reference's own convention,
/root/reference/benches/density.rs:29,48). N>1: one process per GPU (torchrun), every rank owns one 1 GiB shard of ONE
bit-exact stream; the only collective is the all_gather of the 256 KiB dictionary tables (weak scaling).
Rank 0 prints ONE JSON line.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GiB = 1 << 30
METRIC = "chameleon_encode_input_GBps"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bytes", type=int, default=GiB, help="bytes per GPU (default 1 GiB: BASELINE.json configs[1])")
    ap.add_argument("--cpu-sample-bytes", type=int, default=256 << 20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def time_oracle(sample, repeats):
    """Reference algorithm (oracle port, oracle/density_oracle.c) on ONE host core: the reference is single-threaded
    (README.md:42) and one stream cannot be split without changing its bytes."""
    import numpy as np
    import oracle
    L = oracle.lib()
    cap = oracle.safe_encode_buffer_size("chameleon", sample.size)
    out = np.zeros(cap, dtype=np.uint8)      # pre-faulted: page faults are not part of the codec
    L.oracle_encode(0, sample.ctypes.data, min(sample.size, 8 << 20), out.ctypes.data, cap)  # untimed warm-up
    times = []
    for _ in range(repeats):
        t = time.perf_counter()
        n = L.oracle_encode(0, sample.ctypes.data, sample.size, out.ctypes.data, cap)
        times.append(time.perf_counter() - t)
        assert n > 0
    return times, n


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from density_b200 import synth
    nb = min(args.bytes, args.cpu_sample_bytes)
    sample = synth.synth_text(nb).numpy()
    time_oracle(sample[: min(nb, 16 << 20)], max(1, min(args.warmup, 3)))
    times, n = time_oracle(sample, args.steps)
    tot = sum(times)
    val = nb * len(times) / tot / 1e9
    cores = 1
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": "Chameleon encode, synthetic English text (BASELINE.json configs[1])", "bytes_per_step": nb,
                   "note": "reference algorithm restated in C (oracle/density_oracle.c; no Rust toolchain on the box), "
                           "1 thread: the reference is single-threaded and one stream cannot be split bit-exactly"},
        "cpu_baseline": {"value": val, "unit": "GB/s", "cores": cores, "kind": "port",
                         "sample": f"{nb >> 20} MiB prefix of the 1 GiB synthetic-text workload per step"},
        "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "host_cpus": os.cpu_count(), "ratio": sample.size / n,
    }
    print(json.dumps(line))


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import density_b200
    from density_b200 import synth, sharded

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the density_b200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"      # keep rank 0's stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    L = density_b200.load()
    C = density_b200.Chameleon
    n = args.bytes
    assert n % 256 == 0
    peak_gbs, peak_src = measured_peaks()

    # ---- workload: rank r owns pages [r*n/64Ki, ...) of the infinite synthetic corpus: one contiguous stream ----------
    d_in = synth.synth_text(n, device=dev, first_page=rank * (n // synth.PAGE))
    cap = C.safe_encode_buffer_size(n)
    d_out = torch.empty(cap, dtype=torch.uint8, device=dev)
    d_size = torch.zeros(1, dtype=torch.int64, device=dev)
    d_flags = torch.zeros(1, dtype=torch.int32, device=dev)
    enc = sharded.ShardedChameleonEncoder() if world > 1 else None

    def step():
        if world > 1:
            enc.encode(d_in, d_out, d_size, d_flags)
        else:
            density_b200.encode_device("chameleon", d_in, d_out, d_size)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    L.density_b200_profile_enable(1 if world == 1 else 0)
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    out_bytes = int(d_size.item())
    assert out_bytes > 0
    if world == 1:
        assert L.density_b200_last_encode_was_fast() == 1, "synthetic text must take the segment-parallel path"
    else:
        assert int(d_flags.item()) == 0

    # ---- timed region: device-resident, CUDA events on the launching stream, inputs (1 GiB) >> L2 (126 MB) ---------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = L.density_b200_kernel_launches()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    L.density_b200_profile_enable(1 if world == 1 else 0)   # reset the per-stage event ring: it now covers exactly the timed steps
    ev[0].record()
    for i in range(args.steps):
        step()
        ev[i + 1].record()
    barrier()
    prof = (ctypes.c_float * 3)()
    if world == 1 and L.density_b200_profile_get(prof) != 0:
        raise SystemExit("profile_get failed: " + density_b200._lib.last_error())
    total_ms = ev[0].elapsed_time(ev[args.steps])
    launches = L.density_b200_kernel_launches() - launches0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * n / (ms_per_step * 1e-3) / 1e9

    # ---- e2e: the reference-facing symbol chameleon_encode() with HOST (pinned) buffers, copies inside the timing ---
    e2e = None
    if not args.no_e2e:
        h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        h_in.copy_(d_in)
        h_out = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
        a_in, a_out = h_in.numpy(), h_out.numpy()
        h_size = torch.zeros(1, dtype=torch.int64, pin_memory=True)

        def e2e_step():
            if world == 1:
                return C.encode(a_in, a_out)          # reference-shaped C ABI symbol, host pointers
            # N > 1: the public sharded API, same host buffers; H2D / D2H copies are part of the step
            d_in.copy_(h_in, non_blocking=True)
            enc.encode(d_in, d_out, d_size, d_flags)
            h_size.copy_(d_size, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            mm = int(h_size.item())
            h_out[:mm].copy_(d_out[:mm], non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return mm

        for _ in range(3):
            m = e2e_step()
        assert m == out_bytes
        k = max(3, min(args.steps, 10))
        barrier()
        t0 = time.perf_counter()
        for _ in range(k):
            m = e2e_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * n * k / float(tt.item()) / 1e9, "unit": "GB/s", "h2d_bytes_per_step": n, "d2h_bytes_per_step": int(m) + 8,
               "steps": k, "api": ("chameleon_encode(host ptr, n, host ptr, cap): C ABI, pinned host buffers, synchronous" if world == 1 else
                                   "ShardedChameleonEncoder.encode with pinned host buffers: H2D + phase1 + all_gather + phase2 + D2H per step")}

    # ---- extra (not the headline metric): Chameleon decode of the stream just produced, device-resident, N=1 only -------
    extra = None
    if world == 1:
        d_dec = torch.empty(n, dtype=torch.uint8, device=dev)
        d_dsz = torch.zeros(1, dtype=torch.int64, device=dev)
        for _ in range(3):
            density_b200.decode_device("chameleon", d_out, out_bytes, d_dec, d_dsz)
        torch.cuda.synchronize()
        assert int(d_dsz.item()) == n and torch.equal(d_dec, d_in), "decode(encode(x)) != x"
        k = max(3, min(args.steps, 10))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            density_b200.decode_device("chameleon", d_out, out_bytes, d_dec, d_dsz)
        e1.record(); torch.cuda.synchronize()
        dms = e0.elapsed_time(e1) / k
        extra = {"chameleon_decode_GBps": n / (dms * 1e-3) / 1e9, "decode_ms": dms, "round_trip_verified": True,
                 "note": "uncompressed bytes / time, same convention as the reference's decompress bench (benches/density.rs:48)"}
        del d_dec
        # Cheetah / Lion run-parallel encoders on the same buffer (bit-exactness is the tests' job; here: settled copy map, timing)
        for alg in ("cheetah", "lion"):
            C2 = density_b200.CODECS[alg]
            d_o2 = torch.empty(C2.safe_encode_buffer_size(n), dtype=torch.uint8, device=dev)
            d_s2 = torch.zeros(1, dtype=torch.int64, device=dev)
            for _ in range(2):
                density_b200.encode_device(alg, d_in, d_o2, d_s2, path=1)
            torch.cuda.synchronize()
            m2 = int(d_s2.item())
            assert m2 > 0, f"{alg}: copy map did not settle on the bench input"
            e0.record()
            for _ in range(3):
                density_b200.encode_device(alg, d_in, d_o2, d_s2, path=1)
            e1.record(); torch.cuda.synchronize()
            ams = e0.elapsed_time(e1) / 3
            extra[f"{alg}_encode_GBps"] = n / (ams * 1e-3) / 1e9
            extra[f"{alg}_encode_ms"] = ams
            extra[f"{alg}_ratio"] = n / m2
            del d_o2

    # ---- CPU baseline (rank 0, N=1 only) ---------------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        nb = min(n, args.cpu_sample_bytes)
        sample = d_in[:nb].cpu().numpy()
        time_oracle(sample[: 16 << 20], 1)
        reps = 8
        times, _ = time_oracle(sample, reps)
        cpu = {"value": nb * reps / sum(times) / 1e9, "unit": "GB/s", "cores": 1, "kind": "port",
               "sample": f"{nb >> 20} MiB prefix of the step's buffer x {reps} (oracle/density_oracle.c, 1 thread; host has {os.cpu_count()} CPUs)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    line = {
        "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
        "data": "synthetic",
        "config": {"workload": "Chameleon encode, 1 GiB synthetic English text per GPU (BASELINE.json configs[1])",
                   "bytes_per_gpu": n, "out_bytes_rank0": out_bytes, "ratio": n / out_bytes,
                   "l2_policy": "input 1 GiB + output 0.57 GiB per step >> 126 MB L2 (no flush needed)",
                   "parallelism": f"{world} shards of one bit-exact stream; all_gather of 256 KiB tables" if world > 1 else "single GPU"},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if e2e:
        line["e2e"] = e2e
    if world == 1:
        fp, mid, em = [float(x) for x in prof]
        sig = n / 32
        step_alg = (n + out_bytes) / (ms_per_step * 1e-3) / 1e9
        kernels = {
            "cham_flag_pass": {"ms": fp, "alg_bytes": n + sig, "gbs": (n + sig) / (fp * 1e-3) / 1e9 if fp > 0 else None},
            "cham_emit": {"ms": em, "alg_bytes": n + sig + out_bytes, "gbs": (n + sig + out_bytes) / (em * 1e-3) / 1e9 if em > 0 else None},
            "carry_resolve_sizes_scan": {"ms": mid},
        }
        dom = "cham_flag_pass" if fp >= em else "cham_emit"
        ach = kernels[dom]["gbs"]
        traffic = None   # dram__bytes_read.sum + dram__bytes_write.sum of that kernel from the committed ncu --set full capture
        try:
            if n == GiB:
                traffic = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))["dram_bytes_per_launch"][dom]
        except Exception:
            traffic = None
        line["roofline"] = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak_gbs, "unit": "GB/s",
                            "frac": (ach / peak_gbs) if ach else None, "traffic": traffic, "peak_source": peak_src,
                            "step_algorithmic": {"bytes": n + out_bytes, "achieved": step_alg, "frac": step_alg / peak_gbs},
                            "input_rate_frac": value / peak_gbs, "kernels": kernels}
    if extra:
        line["extra"] = extra
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
