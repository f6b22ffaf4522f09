#!/usr/bin/env python
"""bench.py -- Lepton arithmetic-coding hot path on B200: JPEG MB/s, roofline fraction, CPU baseline.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line.

A "step" is one pass of the hot path (encode: coefficient planes -> per-segment bool-coder streams) over one batch of
synthetic input.  The workload is BASELINE.json configs[1]: 4096 x synthetic 1920x1080 4:2:0 q=85 baseline JPEGs per
GPU (weak scaling: every rank codes its own batch, no collective on the data path).
  value   = JPEG MB/s (10^6 input-JPEG bytes per second), kernel time only, planes resident in HBM (CUDA events)
  e2e     = same metric through the file-level C ABI with HOST buffers: JPEG bytes in host memory -> .lep bytes in
            host memory (host Huffman decode + H2D + kernel + D2H + container), wall clock
  --impl reference : the unmodified reference CLI (oracle/_ref/lepton) on the box's host cores, bounded sample.
"""
import argparse
import io
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W_, H_, Q_ = 1920, 1080, 85
WORKLOAD = "4096x synthetic 1920x1080 4:2:0 q=85 baseline JPEGs, encode"


# ---------------------------------------------------------------------------------------------- synthetic corpus
def synth_pixels(seed, w=W_, h=H_):
    """Deterministic photo-like content: low-frequency gradients + band-limited noise (seed = image index)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.zeros((h, w, 3), np.float32)
    for c in range(3):
        a = rng.uniform(-1, 1, 6)
        img[..., c] = 128 + 60 * (a[0] * xx / w + a[1] * yy / h) + 40 * np.sin(
            2 * np.pi * (a[2] * 3 * xx / w + a[3] * 2 * yy / h) + a[4] * 6)
    for scale, amp in ((16, 28.0), (4, 12.0), (1, 4.0)):
        n = rng.normal(0, 1, (h // scale + 2, w // scale + 2)).astype(np.float32)
        if scale > 1:
            n = np.kron(n, np.ones((scale, scale), np.float32))[:h, :w]
            s2 = scale // 2
            n = (n + np.roll(n, s2, 0) + np.roll(n, s2, 1) + np.roll(np.roll(n, s2, 0), s2, 1)) / 4
        else:
            n = n[:h, :w]
        img += amp * n[..., None] * rng.uniform(0.6, 1.0, 3).astype(np.float32)
    return np.clip(img, 0, 255).astype(np.uint8)


def synth_jpeg(seed):
    from PIL import Image
    b = io.BytesIO()
    Image.fromarray(synth_pixels(seed), "RGB").save(b, "JPEG", quality=Q_, subsampling=2, optimize=False)
    return b.getvalue()


def make_corpus(distinct, seed0=0):
    """`distinct` different JPEGs (generated in parallel processes)."""
    from concurrent.futures import ProcessPoolExecutor
    workers = min(distinct, os.cpu_count() or 1, 32)
    with ProcessPoolExecutor(workers) as ex:
        return list(ex.map(synth_jpeg, range(seed0, seed0 + distinct)))


def effective_cores():
    """CPU cores this process may actually use: affinity mask and cgroup quota, whichever is smaller."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q, p = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if q != "max":
            n = min(n, max(1, int(float(q) / float(p))))
    except Exception:
        pass
    return max(1, n)


# ---------------------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.samples = []
        self.reasons = set()
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                self.samples.append((float(f[0]), float(f[1])))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    self.reasons.add(name)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = sorted(s[0] for s in self.samples)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.samples[0][1] if self.samples else None,
                "reasons": sorted(self.reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- reference arm / cpu baseline
REF_LEPTON = os.path.join(ROOT, "oracle", "_ref", "lepton")


def run_reference_sample(jpegs, workers):
    """Encode every JPEG with the UNMODIFIED reference CLI, `workers` concurrent processes ("backfill" mode of
    src/lepton/benchmark.cc:388-423).  Returns (seconds, bytes)."""
    from concurrent.futures import ThreadPoolExecutor
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as td:
        paths = []
        for i, j in enumerate(jpegs):
            p = os.path.join(td, "i%05d.jpg" % i)
            with open(p, "wb") as f:
                f.write(j)
            paths.append(p)

        def one(p):
            r = subprocess.run([REF_LEPTON, "-skipverify", "-unjailed", p, p[:-4] + ".lep"], stdout=subprocess.DEVNULL,
                               stderr=subprocess.DEVNULL)
            return r.returncode
        t0 = time.perf_counter()
        with ThreadPoolExecutor(workers) as ex:
            rcs = list(ex.map(one, paths))
        dt = time.perf_counter() - t0
        if any(rcs):
            raise RuntimeError("reference CLI failed on the sample: %r" % rcs[:8])
    return dt, sum(len(j) for j in jpegs)


def cpu_baseline(distinct_jpegs, sample_files):
    cores = effective_cores()
    jp = [distinct_jpegs[i % len(distinct_jpegs)] for i in range(sample_files)]
    dt, nbytes = run_reference_sample(jp, cores)
    return {"value": nbytes / dt / 1e6, "unit": "MB/s", "cores": cores, "kind": "reference",
            "sample": "%d files (%.1f MB JPEG) of the workload through oracle/_ref/lepton -skipverify -unjailed, %d concurrent "
                      "processes, wall clock incl. process spawn" % (sample_files, nbytes / 1e6, cores)}


# ---------------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--images", type=int, default=4096, help="images per GPU per step")
    ap.add_argument("--distinct", type=int, default=32, help="distinct synthetic images replicated to --images")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-decode", action="store_true")
    ap.add_argument("--host-threads", type=int, default=0, help="host threads for the e2e stage (0 = effective cores / ranks)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=512)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != max(args.gpus, 1) and world > 1:
        args.gpus = world

    config = {"workload": WORKLOAD, "images_per_gpu": args.images, "distinct_images": args.distinct, "width": W_, "height": H_,
              "subsampling": "4:2:0", "quality": Q_, "l2": "inputs (%.1f GB of coefficient planes per GPU) far exceed the 126 MB L2"
                                                            % (args.images * 6266880 / 1e9),
              "parallelism": "independent per-GPU batches x%d (no collective on the data path)" % max(world, 1)}

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        if not os.path.exists(REF_LEPTON):
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/lepton not built (run __graft_entry__.build() where /root/reference exists)"}))
            return 0
        distinct = make_corpus(min(args.distinct, 32))
        cores = effective_cores()
        sample = [distinct[i % len(distinct)] for i in range(args.cpu_sample)]
        for _ in range(max(args.warmup, 0)):
            run_reference_sample(sample[:max(8, cores // 4)], cores)
        tot_t, tot_b = 0.0, 0
        for _ in range(args.steps):
            dt, nb = run_reference_sample(sample, cores)
            tot_t += dt
            tot_b += nb
        v = tot_b / tot_t / 1e6
        line = {"impl": "reference", "metric": "JPEG MB/s encode", "value": v, "unit": "MB/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "int16 coefficients / u8 probabilities (integer)",
                "data": "synthetic", "config": config,
                "cpu_baseline": {"value": v, "unit": "MB/s", "cores": cores, "kind": "reference",
                                 "sample": "%d files per step through oracle/_ref/lepton -skipverify -unjailed, %d concurrent processes"
                                           % (len(sample), cores)},
                "e2e": {"value": v, "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm
    import numpy as np
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from lepton_b200 import HostJpeg, LeptonB200Codec, LeptonB200FileCodec

    distinct = make_corpus(args.distinct, seed0=1000 * rank)
    jpegs = [distinct[i % len(distinct)] for i in range(args.images)]
    jpeg_bytes = sum(len(j) for j in jpegs)

    # host front end once (outside the timed region) for the device-resident measurement
    hjs = [HostJpeg(j) for j in distinct]
    for h in hjs:
        assert h.status == 0, h.error
    base_imgs = [h.coef_image() for h in hjs]
    imgs = [base_imgs[i % len(base_imgs)] for i in range(args.images)]
    nseg = sum(im.nseg for im in imgs)
    blocks = sum(im.blocks() for im in imgs)

    codec = LeptonB200Codec(local_rank)
    codec.encode_upload(imgs)            # H2D: planes become resident in HBM
    codec.encode_launch()
    res = codec.encode_fetch(copy=False)
    assert all(s.status == 0 for r in res for s in r), "encode failed"
    stream_bytes = sum(codec.last_lens)
    ndecisions = sum(s.ndecisions for r in res for s in r)
    alg_bytes = codec.last_algorithmic_bytes

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup - 1, 0)):
        codec.encode_launch()
        codec.sync()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = codec.kernel_launches
    t_wall0 = time.perf_counter()
    kernel_ms, sym_ms = [], []
    for _ in range(args.steps):
        codec.encode_launch()
        codec.sync()
        kernel_ms.append(codec.last_kernel_ms)
        sym_ms.append(codec.last_symbolise_ms)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = codec.kernel_launches - launches0
    clocks = sampler.stop()
    from lepton_b200.sharding import reduce_job_throughput
    dev_s = sum(kernel_ms) / 1e3
    # whole-job throughput: bytes of ALL ranks / max over ranks of the device time (same helper the gloo test covers)
    thr, total_units, dev_s_max = reduce_job_throughput(jpeg_bytes * args.steps, dev_s, dist, "cuda")
    _, _, wall_max = reduce_job_throughput(0.0, t_wall, dist, "cuda")
    total_jpeg = total_units / args.steps
    value = thr / 1e6

    # ---------------------------------------------------------------- decode direction + round trip (same batch)
    decode = None
    if not args.no_decode:
        res = codec.encode_fetch(copy=True)                       # streams of the last launch, on the host
        streams = [[s.data for s in r] for r in res]
        # round trip on the distinct images: GPU decode of the GPU-coded streams must give back the input planes
        from lepton_b200 import CoefImage
        outs = [CoefImage(ncmp=im.ncmp, mcuv=im.mcuv, bch=im.bch, bcv=im.bcv, qtables_zigzag=im.qtables_zigzag,
                          planes=[np.full_like(p, 1) for p in im.planes], luma_y_start=im.luma_y_start) for im in base_imgs]
        st = codec.decode_images(outs, streams[:len(outs)])
        ok = 0
        k = 0
        for im, o in zip(base_imgs, outs):
            good = all(s == 0 for s in st[k:k + im.nseg]) and all(np.array_equal(a, b) for a, b in zip(im.planes, o.planes))
            ok += int(good)
            k += im.nseg
        # timing: whole batch, streams resident in HBM
        codec.decode_upload(imgs, streams)
        for _ in range(max(args.warmup - 1, 1)):
            codec.decode_launch()
            codec.sync()
        barrier()
        dms = []
        for _ in range(args.steps):
            codec.decode_launch()
            codec.sync()
            dms.append(codec.last_kernel_ms)
        barrier()
        td = torch.tensor([sum(dms) / 1e3], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(td, op=dist.ReduceOp.MAX)
        decode = {"value": jpeg_bytes * max(world, 1) * args.steps / float(td[0]) / 1e6, "unit": "MB/s", "ms_per_step": 1e3 * float(td[0]) / args.steps,
                  "kernel": "lep_decode_kernel", "roundtrip_pass_rate": ok / len(base_imgs), "roundtrip_images": len(base_imgs),
                  "decisions_per_s": ndecisions / (sum(dms) / len(dms) / 1e3)}

    # ---------------------------------------------------------------- e2e through the file-level C ABI (host buffers)
    e2e = None
    if not args.no_e2e:
        codec.close()
        codec = None
        threads = args.host_threads or max(1, effective_cores() // max(world, 1))
        fc = LeptonB200FileCodec(local_rank, host_threads=threads)
        handle = fc.prepare(jpegs)                  # pointer/length array of the host buffers (ctypes marshalling, once)
        r = fc.compress(handle, copy=True)          # warm-up (allocates pinned arenas); keep the .lep files for the way back
        assert all(st == 0 for st, _ in r)
        leps = [b for _, b in r]
        lep_bytes = sum(len(b) for b in leps)
        barrier()
        l0 = fc.kernel_launches
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            fc.compress(handle, copy=False)
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": total_jpeg * args.e2e_steps / float(tt[0]) / 1e6, "unit": "MB/s",
               "h2d_bytes_per_step": int(jpeg_bytes), "d2h_bytes_per_step": int(stream_bytes),
               "h2d_note": "entropy-coded scan bytes (Huffman decode happens on the GPU); files the host has to decode upload 128 B per block instead",
               "steps": args.e2e_steps, "host_threads": threads, "api": "lepb200_compress_jpegs (JPEG bytes -> .lep bytes, host memory)",
               "stage_seconds_last_step": fc.last_timing(), "lep_bytes_per_step": int(lep_bytes),
               "gpu_launches": fc.kernel_launches - l0}
        # the way back through the same API: .lep bytes -> JPEG bytes (GPU arithmetic decode, host Huffman re-encode)
        if not args.no_decode:
            def all_ok(ok):                  # ranks agree before any collective follows (a failed rank must not leave the others waiting)
                t = torch.tensor([0 if ok else 1], dtype=torch.int32, device="cuda")
                if dist is not None:
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                return int(t[0]) == 0
            err, exact, dt2 = None, 0, None
            try:
                lhandle = fc.prepare(leps)
                back = fc.decompress(lhandle, copy=True)       # warm-up + round-trip check of every file
                exact = sum(int(st == 0 and b == j) for (st, b), j in zip(back, jpegs))
                del back
            except Exception as ex:          # e.g. pinned host memory for the plane arenas not available on this box
                err = str(ex)[:200]
            if all_ok(err is None):
                barrier()
                t0 = time.perf_counter()
                try:
                    fc.decompress(lhandle, copy=False)
                except Exception as ex:
                    err = str(ex)[:200]
                barrier()
                dt2 = time.perf_counter() - t0
            if all_ok(err is None) and dt2 is not None:
                td2 = torch.tensor([dt2], dtype=torch.float64, device="cuda")
                if dist is not None:
                    dist.all_reduce(td2, op=dist.ReduceOp.MAX)
                e2e["decode"] = {"value": total_jpeg / float(td2[0]) / 1e6, "unit": "MB/s", "steps": 1,
                                 "api": "lepb200_decompress_leps (.lep bytes -> JPEG bytes, host memory)",
                                 "roundtrip_pass_rate": exact / len(jpegs), "roundtrip_files": len(jpegs),
                                 "stage_seconds": fc.last_timing()}
            else:
                e2e["decode"] = {"value": None, "unit": "MB/s", "error": err or "failed on another rank"}
        fc.close()

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("hbm_gbs", 6650.0)
    avg_step_s = (sum(kernel_ms) / len(kernel_ms)) / 1e3          # kernel A + kernel B
    avg_launch_s = (sum(sym_ms) / len(sym_ms)) / 1e3               # dominant kernel: A (symbolise + model update)
    achieved = alg_bytes / avg_launch_s / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "encode_kernel_traffic.json"))).get("dram_bytes_per_launch_4096")
    except Exception:
        pass
    line = {
        "metric": "JPEG MB/s encode", "value": value, "unit": "MB/s", "n_gpus": max(world, 1), "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dev_s_max / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int16 coefficients / u8 probabilities (integer)", "data": "synthetic",
        "config": config, "clocks": clocks, "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                     "algorithmic_bytes_per_launch": int(alg_bytes), "kernel": "lep_encode_kernel (symbolise + model update); lep_rangecode_kernel is the remainder of the step",
                     "kernel_ms": 1e3 * avg_launch_s, "rangecode_kernel_ms": 1e3 * (avg_step_s - avg_launch_s),
                     "step_frac": alg_bytes / avg_step_s / 1e9 / peak,
                     "decisions_per_s": ndecisions / avg_step_s, "decisions_per_launch": int(ndecisions)},
        "wall_ms_per_step": 1e3 * wall_max / args.steps,
        "batch": {"jpeg_bytes": int(jpeg_bytes), "segments": int(nseg), "blocks": int(blocks), "stream_bytes": int(stream_bytes)},
    }
    if decode:
        line["decode"] = decode
        line["roundtrip_pass_rate"] = decode["roundtrip_pass_rate"]       # BASELINE.json: "bit-exact round-trip pass rate"
    if e2e:
        line["e2e"] = e2e
    if not args.no_cpu_baseline and os.path.exists(REF_LEPTON):
        line["cpu_baseline"] = cpu_baseline(distinct, args.cpu_sample)
    elif not args.no_cpu_baseline:
        line["cpu_baseline"] = {"value": None, "unit": "MB/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref/lepton missing"}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
