#!/usr/bin/env python3
"""bench.py -- decompressed GB/s of the B200 LZ4 path on BASELINE.json's headline workload.

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA kernels via the C-ABI)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port)

Workload at every N (weak scaling: fixed work per GPU): BASELINE.json configs[1] -- a synthetic
4 GiB corpus of text-like data (ratio ~2.2), LZ4 frames with 64 KiB independent blocks, block and
content XXH32, organised as 4096 frames x 1 MiB so that content-checksum chains run concurrently.
One step = one pass of the hot path over the whole corpus.

  value     decompressed bytes / s with the compressed corpus already resident in HBM
            (K1 decode + K3 content checksums + status D2H + host fold), all ranks, max-over-ranks time
  e2e       the same metric through the public batch call with HOST buffers: block-table build,
            H2D of the compressed bytes, kernels, D2H of the decompressed bytes, every step
  roofline  K1 (the dominant kernel): algorithmic bytes (compressed read + decompressed written)
            / its CUDA-event duration, against the measured HBM copy peak (MEASURED_PEAKS.json)
  cpu_baseline  the oracle (C restatement of lib/lz4ada.adb) on this box's host cores, bounded sample

No CPU fallback: if the CUDA library cannot be loaded or there is no device, our arm exits non-zero.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

GIB = 1 << 30


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size-gib", type=float, default=4.0, help="decompressed bytes per GPU")
    ap.add_argument("--frame-mib", type=float, default=1.0)
    ap.add_argument("--block", default="64k", choices=["64k", "256k", "1m", "4m"])
    ap.add_argument("--kinds", default="text", help="comma list of text,rle,random")
    ap.add_argument("--no-block-checksum", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-sample-mib", type=int, default=0, help="0 = auto")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--k1-group", type=int, default=0,
                    help="K1 tuning (include/lz4b200.h): 0 auto, 50 v5, 40 v4, 64 v3, 1..16 v2, -1 v1")
    return ap.parse_args()


K1_PROFILE = "r01_s2_k1_ncu_full_4gib.csv"   # ncu --set full of the headline workload, summarised by tools/ncu_summary.py


def ncu_traffic(args, kernel):
    """DRAM bytes (read + write) of one K1 launch from the committed `ncu --set full` capture of this
    exact workload and kernel (profiles/K1_PROFILE); None for any other workload or kernel."""
    if not (args.size_gib == 4.0 and args.kinds == "text" and args.block == "64k" and args.frame_mib == 1.0
            and not args.no_block_checksum):
        return None
    try:
        rd = wr = None
        name = ""
        with open(os.path.join(ROOT, "profiles", K1_PROFILE)) as f:
            for line in f:
                parts = line.strip().split(",")
                if parts[0] == "Kernel Name":
                    name = parts[2]
                if parts[0] == "dram__bytes_read.sum":
                    rd = float(parts[2]) * {"Gbyte": 1e9, "Mbyte": 1e6}[parts[1]]
                if parts[0] == "dram__bytes_write.sum":
                    wr = float(parts[2]) * {"Gbyte": 1e9, "Mbyte": 1e6}[parts[1]]
        if kernel not in name:
            return None
        return int(rd + wr) if rd is not None and wr is not None else None
    except Exception:
        return None


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons during the timed region (pynvml; nvidia-smi fields equivalent)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def make_corpus(args, rank):
    from tools import corpus
    code = {"64k": 4, "256k": 5, "1m": 6, "4m": 7}[args.block]
    frame_bytes = int(args.frame_mib * (1 << 20))
    total = int(args.size_gib * GIB)
    kinds = tuple(args.kinds.split(","))
    t0 = time.time()
    c = corpus.build_corpus(total, frame_bytes, code, kinds=kinds, block_checksum=not args.no_block_checksum,
                            content_checksum=True, seed=1234 + 1000003 * rank,
                            workers=max(4, (os.cpu_count() or 8) // max(1, args.gpus)))
    c["build_s"] = time.time() - t0
    c["frame_bytes"] = frame_bytes
    return c


def workload_name(args):
    return ("synthetic %.3g GiB per GPU, %s, LZ4 frames of %.3g MiB, %s independent blocks, %s+content XXH32"
            % (args.size_gib, "+".join(args.kinds.split(",")), args.frame_mib, args.block,
               "content" if args.no_block_checksum else "block"))


# --------------------------------------------------------------------------------------- CPU arm
def cpu_decode_throughput(c, frames, threads):
    """Oracle (C restatement of lib/lz4ada.adb, Init(For_All)+Update fed 4 KiB like unlz4ada_simple)
    over `frames` frame indices using `threads` host threads.  -> (GB/s, seconds, bytes)"""
    import oracle_binding
    o = oracle_binding.load()
    u8p = ctypes.POINTER(ctypes.c_uint8)
    src = np.frombuffer(c["src"], dtype=np.uint8)
    base = src.ctypes.data
    fb = c["frame_bytes"]

    def work(idxs):
        out = np.empty(fb + 64, dtype=np.uint8)
        n, eof, msg = ctypes.c_size_t(0), ctypes.c_int(0), ctypes.create_string_buffer(700)
        done = 0
        for i in idxs:
            off, ln = c["items"][i]
            rc = o.lib.lzo_decode_stream(ctypes.cast(base + off, u8p), ln, 4096, out.ctypes.data_as(u8p), fb + 64,
                                         ctypes.byref(n), ctypes.byref(eof), msg, 700)
            assert rc == 0 and n.value == fb, msg.value
            done += n.value
        return done

    parts = [frames[k::threads] for k in range(threads)]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        total = sum(ex.map(work, parts))
    dt = time.perf_counter() - t0
    return total / dt / 1e9, dt, total


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # bounded sample: sized so that warmup + steps finish in a few minutes
    sample_mib = args.cpu_sample_mib or min(int(args.size_gib * 1024), 64 * cores)
    small = argparse.Namespace(**vars(args))
    small.size_gib = sample_mib / 1024.0
    c = make_corpus(small, 0)
    frames = list(range(len(c["items"])))
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_decode_throughput(c, frames, cores)
    times, total = [], 0
    for _ in range(args.steps):
        g, dt, total = cpu_decode_throughput(c, frames, cores)
        times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = total / (ms / 1e3) / 1e9
    print(json.dumps({
        "impl": "reference", "metric": "decompressed_GBps", "value": value, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name(args), "sample": "%d MiB of the workload per step" % sample_mib},
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": cores, "kind": "port",
                         "sample": "%d MiB (%d frames) per step, one frame per thread task, oracle = C restatement "
                                   "of lib/lz4ada.adb (no GNAT in this image)" % (sample_mib, len(frames))},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# --------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import bo_lz4_ada_b200 as lz

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        sys.exit("bench.py: no CUDA device -- this arm has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the context launches on torch's current stream so that torch.cuda.Event sees the kernels
    stream = torch.cuda.current_stream()
    ctx = lz.DeviceContext(local, stream.cuda_stream)
    ctx.set_tuning(args.k1_group)
    c = make_corpus(args, rank)
    src_np = np.frombuffer(c["src"], dtype=np.uint8)
    n_src = len(src_np)

    # pinned host buffers (the e2e leg copies from / to these every step)
    h_src = torch.empty(n_src + 64, dtype=torch.uint8).pin_memory()
    h_src[:n_src].copy_(torch.from_numpy(src_np.copy()))
    batch = lz.Batch(ctx, h_src.data_ptr(), c["items"])
    batch.src_bytes = n_src
    out_bytes = batch.output_bytes
    d_src = torch.empty(n_src + 256, dtype=torch.uint8, device="cuda")
    d_dst = torch.empty(out_bytes + 256, dtype=torch.uint8, device="cuda")
    batch.upload(d_src.data_ptr())
    torch.cuda.synchronize()

    def step():
        batch.run(d_src.data_ptr(), d_dst.data_ptr())

    for _ in range(max(args.warmup, 3)):
        step()
    res = batch.results()
    bad = [r for r in res if r["exception"] != "OK"]
    assert not bad, bad[:2]
    plain_bytes = sum(r["out_len"] for r in res)
    assert plain_bytes == c["plain_bytes"], (plain_bytes, c["plain_bytes"])
    # every frame's content checksum was recomputed on the device and compared with the encoder's

    sampler = ClockSampler(local)
    sampler.start()
    launches0 = ctx.launch_count()
    k1_ms, k3_ms = [], []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
        km = batch.kernel_ms()
        k1_ms.append(km["k1_decode_blocks"])
        k3_ms.append(km["k3_xxh32_frames"])
    e1.record()
    barrier()
    sampler.stop_flag = True
    elapsed_ms = e0.elapsed_time(e1)
    launches = ctx.launch_count() - launches0
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device="cuda")
    total_plain = torch.tensor([float(plain_bytes)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(total_plain, op=dist.ReduceOp.SUM)
    ms_per_step = float(t.item()) / args.steps
    value = float(total_plain.item()) / (ms_per_step / 1e3) / 1e9

    # ---- e2e: public batch call on host buffers, H2D + kernels + D2H inside the timed region
    e2e = None
    if not args.skip_e2e:
        h_dst = torch.empty(out_bytes + 64, dtype=torch.uint8).pin_memory()
        items = (lz.BatchItem * len(c["items"]))()
        results = (lz.BatchResult * len(c["items"]))()

        def e2e_step():
            for k, (off, ln) in enumerate(c["items"]):
                items[k].src_off, items[k].src_len, items[k].dst_off, items[k].dst_cap = off, ln, 0, 0
            rc = lz.lib().lz4ada_batch_decompress(ctx.handle, h_src.data_ptr(), n_src, h_dst.data_ptr(), out_bytes,
                                                  len(c["items"]), items, lz.RESERVATIONS["For_All"], results, None, 0)
            assert rc == 0, rc

        e2e_step()
        assert all(results[k].exception == 0 for k in range(len(c["items"])))
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / args.e2e_steps
        te = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": float(total_plain.item()) / float(te.item()) / 1e9, "unit": "GB/s",
               "h2d_bytes_per_step": int(n_src), "d2h_bytes_per_step": int(plain_bytes),
               "ms_per_step": 1e3 * float(te.item()),
               "note": "lz4ada_batch_decompress on pinned host buffers: block-table build + H2D + K1/K3 + D2H"}
        # spot-check the bytes that came back against the encoder-side digests
        from tools import corpus as _c
        for k in (0, len(c["items"]) // 2, len(c["items"]) - 1):
            got = bytes(h_dst[results[k].dst_off:results[k].dst_off + results[k].out_len].numpy())
            assert _c.xxh32(got) == c["digests"][k]

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    traffic = batch.traffic()
    peak, peak_src = peaks()
    k1_name = batch.k1_kernel_name() or ctx.k1_kernel_name(batch.block_count)
    k1 = float(np.mean(k1_ms)) if k1_ms else 0.0
    k1_bytes = traffic["compressed_read"] + traffic["decompressed_written"]
    achieved = k1_bytes / (k1 / 1e3) / 1e9 if k1 > 0 else 0.0
    line = {
        "metric": "decompressed_GBps", "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name(args), "frames_per_gpu": len(c["items"]),
                   "blocks_per_gpu": int(batch.block_count), "compressed_bytes_per_gpu": int(n_src),
                   "ratio": c["plain_bytes"] / n_src, "encoder": c["encoder"],
                   "l2": "inputs (%.2f GB) and outputs (%.2f GB) per step exceed the 126 MB L2; no flush needed"
                         % (n_src / 1e9, plain_bytes / 1e9),
                   "corpus_build_s": round(c["build_s"], 1)},
        "clocks": sampler.summary(),
        "gpu_launches": int(launches),
        "kernel_ms": {"k1_decode_blocks": k1, "k3_xxh32_frames": float(np.mean(k3_ms)) if k3_ms else 0.0},
        "roofline": {"bound": "hbm", "kernel": k1_name + " (K1)", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(args, k1_name), "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": int(k1_bytes),
                     "whole_step_frac": (k1_bytes + traffic["checksum_reread"]) / (ms_per_step / 1e3) / 1e9 / peak},
    }
    if e2e:
        line["e2e"] = e2e
    if not args.skip_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        n_frames = len(c["items"])
        sample = list(range(min(n_frames, max(8, 48))))
        g1, dt1, b1 = cpu_decode_throughput(c, sample, 1)
        many = list(range(min(n_frames, 64 * cores)))
        gN, dtN, bN = cpu_decode_throughput(c, many, cores)
        line["cpu_baseline"] = {
            "value": gN, "unit": "GB/s", "cores": cores, "kind": "port",
            "sample": "%d frames (%d MiB) of the same corpus, one frame per task on %d threads; single thread: "
                      "%.3f GB/s on %d frames" % (len(many), bN >> 20, cores, g1, len(sample)),
            "single_thread": {"value": g1, "unit": "GB/s", "cores": 1},
            "note": "oracle = C restatement of lib/lz4ada.adb fed 4 KiB chunks (no GNAT in this image); the "
                    "reference README quotes ~1.1 GB/s (text) single-thread on a Xeon W-2295"}
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
