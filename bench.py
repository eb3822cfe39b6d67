#!/usr/bin/env python
"""bench.py -- encode / decode megapixels per second of the SPIHT hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, sm_100a), BASELINE.json configs[1]
    python bench.py --config {1..5} ...                       # the other BASELINE configurations
    python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on the host cores

Workloads (BASELINE.json `configs`):
  1  images/zebra.jpg (3x256x384), bior2.2 reflect, SpihtSettings() defaults, 1.0 bpp, encode + decode
  2  batch of 256 synthetic 1024x1024 RGB images per GPU, bior2.2 reflect, 0.5 bpp           (headline, default)
  3  1024 synthetic 2048x2048 images over the GPUs (512 on one), IPT, scales [50,15,15], q = 1, 0.1 bpp
  4  4 synthetic 8192x8192 images per GPU, bior6.8 reflect, max level, 1.0 bpp
  5  mixed sizes 512..4096 in one batch, bpp 0.075/0.1/0.5/1.0 x {reflect, periodization}, encode + decode
A "step" is one pass of the hot path (RGB(->IPT) -> DWT -> int32 quantise -> pyramid -> SPIHT encode) over the
batch.  With N > 1 every rank owns its own batch (images are independent: weak scaling, no collective on the data
path; an NCCL all-gather of the per-image stream lengths closes every step).

One JSON line is printed by rank 0 (see the task contract): `value` is whole-job encode MP/s with the inputs
resident in HBM, `e2e` the same metric through the public API with pinned HOST buffers in and host bytes out,
`roofline` the achieved HBM GB/s of the dominant kernel against MEASURED_PEAKS.json, `cpu_baseline` the CPU oracle
(a port of the reference algorithm) timed on the host cores.
"""
import argparse
import json
import multiprocessing as mp
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FALLBACK_HBM_GBS = 6650.0   # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent

# ----------------------------------------------------------------------------- workloads
CONFIGS = {
    1: dict(label="BASELINE.json configs[0]", image="tests/golden/images/zebra.jpg", batch=1, h=256, w=384, bpp=1.0,
            wavelet="bior2.2", mode="reflect", settings={}),
    2: dict(label="BASELINE.json configs[1]", batch=256, h=1024, w=1024, bpp=0.5, wavelet="bior2.2", mode="reflect",
            settings={}),
    3: dict(label="BASELINE.json configs[2]", batch=512, total=1024, h=2048, w=2048, bpp=0.1, wavelet="bior2.2",
            mode="reflect", settings=dict(quantization_scale=1.0, color_model="IPT",
                                          per_channel_quant_scales=[50.0, 15.0, 15.0])),
    4: dict(label="BASELINE.json configs[3]", batch=4, h=8192, w=8192, bpp=1.0, wavelet="bior6.8", mode="reflect",
            settings={}),
    5: dict(label="BASELINE.json configs[4]", mixed=[(512, 64), (1024, 32), (2048, 8), (4096, 2)],
            bpps=[0.075, 0.1, 0.5, 1.0], modes=["reflect", "periodization"], wavelet="bior2.2", settings={}),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help="images per GPU (0 = the configuration's)")
    ap.add_argument("--size", type=int, default=0, help="square image size (0 = the configuration's)")
    ap.add_argument("--bpp", type=float, default=0.0)
    ap.add_argument("--wavelet", default="")
    ap.add_argument("--mode", default="")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-images", type=int, default=0, help="images in the CPU baseline sample (0 = automatic)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-decode", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def resolve(args):
    """the workload as a plain dict (picklable: the CPU workers get it too)"""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    c = dict(CONFIGS[args.config])
    c["config"] = args.config
    if "mixed" in c:
        return c
    if args.size:
        c["h"] = c["w"] = args.size
        c.pop("image", None)
    if args.bpp:
        c["bpp"] = args.bpp
    if args.wavelet:
        c["wavelet"] = args.wavelet
    if args.mode:
        c["mode"] = args.mode
    if c.get("total") and world > 1:
        c["batch"] = max(1, min(c["batch"], c["total"] // world))
    if args.batch:
        c["batch"] = args.batch
    c["max_bits"] = int(c["h"] * c["w"] * c["bpp"])
    return c


def workload_name(c):
    if "mixed" in c:
        sizes = ", ".join(f"{n}x {s}x{s}" for s, n in c["mixed"])
        return (f"mixed-size batch per GPU ({sizes}), {c['wavelet']}, modes {'/'.join(c['modes'])}, "
                f"bpp sweep {'/'.join(str(b) for b in c['bpps'])}, encode + decode ({c['label']})")
    st = c["settings"]
    extra = ""
    if st.get("color_model"):
        extra = f", {st['color_model']} colour space, scales {st['per_channel_quant_scales']}, q = {st['quantization_scale']:g}"
    src = c.get("image") or f"batch of {c['batch']} synthetic {c['h']}x{c['w']} RGB images per GPU"
    return f"{src}, {c['wavelet']} {c['mode']}, {c['bpp']:g} bpp{extra} ({c['label']})"


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons during the timed region"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except ValueError:
                continue
            for name, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- CPU baseline (oracle port)
def _synth_numpy(h, w, seed):
    import numpy as np
    rng = np.random.default_rng(seed)
    fy = np.fft.fftfreq(h)[:, None]
    fx = np.fft.fftfreq(w)[None, :]
    f = np.sqrt(fy * fy + fx * fx)
    f[0, 0] = 1.0
    fields = np.fft.ifft2(np.fft.fft2(rng.normal(size=(4, h, w))) / f).real
    img = 0.8 * fields[:1] + 0.2 * fields[1:]
    img = (img - img.min(axis=(1, 2), keepdims=True)) / np.ptp(img, axis=(1, 2), keepdims=True)
    return img.astype(np.float32).astype(np.float64)


def _cpu_worker(job):
    idx, c, seed = job
    sys.path.insert(0, ROOT)
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from oracle import spiht_oracle, wrapper_ref
    h, w = c["h"], c["w"]
    if c.get("image"):
        from spiht_b200.utils import imload
        img = imload(os.path.join(ROOT, c["image"]))
    else:
        img = _synth_numpy(h, w, seed + idx)
    st = c["settings"]
    kw = dict(wavelet=c["wavelet"], mode=c["mode"], quantization_scale=st.get("quantization_scale", 50.0),
              color_model=st.get("color_model"), per_channel_quant_scales=st.get("per_channel_quant_scales"))
    t0 = time.perf_counter()
    arr, ll_h, ll_w = wrapper_ref.forward_coeffs(img, fast=True, **kw)   # compiled transform (oracle/dwt_fast.c)
    t1 = time.perf_counter()
    data, max_n = spiht_oracle.encode(arr, ll_h, ll_w, c["max_bits"])
    t2 = time.perf_counter()
    rec = spiht_oracle.decode(data, max_n, 3, arr.shape[1], arr.shape[2], ll_h, ll_w)
    t3 = time.perf_counter()
    wrapper_ref.inverse_coeffs(rec, h, w, fast=True, **kw)
    t4 = time.perf_counter()
    return (t1 - t0, t2 - t1, t3 - t2, t4 - t3)


def cpu_reference_sample(c, n_images, cores, seed=4242):
    """the oracle (C restatements of the Rust coder and of the PyWavelets transform, float64; the colour transform
    in numpy) on n_images images of workload `c`, one image per process over `cores` processes.
    Returns (encode MP/s, decode MP/s, detail dict).  Throughput = pixels / (summed per-image codec time / processes
    in use): an ideal-parallel figure that leaves out pool start-up, image synthesis and load imbalance -- it
    flatters the CPU slightly."""
    from oracle import dwt_fast, spiht_oracle
    spiht_oracle.build()
    dwt_fast.lib()
    jobs = [(i, c, seed) for i in range(n_images)]
    tiny = dict(c, h=64, w=64, max_bits=2048, image=None)
    with mp.get_context("spawn").Pool(cores) as pool:
        pool.map(_cpu_worker, [(0, tiny, seed)] * cores)   # warm the workers (imports, build)
        t0 = time.perf_counter()
        res = pool.map(_cpu_worker, jobs, chunksize=1)
        wall = time.perf_counter() - t0
    mp_total = n_images * c["h"] * c["w"] / 1e6
    enc = sum(r[0] + r[1] for r in res)
    dec = sum(r[2] + r[3] for r in res)
    enc_mps = mp_total / (enc / min(cores, n_images))
    dec_mps = mp_total / (dec / min(cores, n_images))
    detail = {"images": n_images, "wall_s": round(wall, 2),
              "per_image_s": {"dwt_quant": round(statistics.mean(r[0] for r in res), 3),
                              "spiht_encode": round(statistics.mean(r[1] for r in res), 3),
                              "spiht_decode": round(statistics.mean(r[2] for r in res), 3),
                              "inverse_dwt": round(statistics.mean(r[3] for r in res), 3)}}
    return enc_mps, dec_mps, detail


def cpu_sample_size(c, cores, requested=0):
    """about 10-30 s of CPU work: images per sample by image area (a 1024^2 image takes ~0.4 s per core)"""
    if requested:
        return requested
    area = c["h"] * c["w"] / (1024 * 1024)
    per_core = max(1, min(4, int(round(1.6 / max(area, 0.05)))))
    n = per_core * cores
    if area >= 16:
        n = max(1, min(cores, 4))       # 8192^2: ~1.7 GB and tens of seconds per image
    return n


CPU_SAMPLE_NOTE = ("oracle/spiht_ref.c (restated Rust coder) + oracle/dwt_fast.c (restated PyWavelets transform, float64)"
                   "; ideal-parallel: pixels / (summed per-image codec time / processes), pool and imbalance cost left out")


def _single_shape_configs(c):
    """config 5 as a list of single-shape workloads (for the CPU arm)"""
    out = []
    for mode in c["modes"]:
        for s, _ in c["mixed"]:
            out.append(dict(c, h=s, w=s, mode=mode, bpp=0.5, max_bits=int(s * s * 0.5)))
    return out


def _cpu_mixed(c, cores, seed=4242):
    """one image of every size and mode of config 5 at 0.5 bpp; throughput over their summed pixels"""
    subs = _single_shape_configs(c)
    pix = sum(s["h"] * s["w"] for s in subs) / 1e6
    te = td = 0.0
    detail = None
    for s in subs:
        e, d, detail = cpu_reference_sample(s, 1, 1, seed=seed)
        te += s["h"] * s["w"] / 1e6 / e
        td += s["h"] * s["w"] / 1e6 / d
    par = min(cores, len(subs))
    return pix / te * par, pix / td * par, detail, len(subs), par


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    c = resolve(args)
    cores = os.cpu_count() or 1
    used = cores
    vals, dvals = [], []
    t_all = time.perf_counter()
    detail = None
    n_images = 0
    for step in range(args.warmup + args.steps):
        if "mixed" in c:
            enc_mps, dec_mps, detail, n_images, used = _cpu_mixed(c, cores, seed=4242 + 100 * step)
        else:
            n_images = cpu_sample_size(c, cores, args.cpu_images)
            enc_mps, dec_mps, detail = cpu_reference_sample(c, n_images, cores, seed=4242 + 100 * step)
            used = min(cores, n_images)
        if step >= args.warmup:
            vals.append(enc_mps)
            dvals.append(dec_mps)
        if time.perf_counter() - t_all > 240 and len(vals) >= 1:
            break
    value = statistics.mean(vals)
    pixels = n_images * (c.get("h", 1024) * c.get("w", 1024))
    bpp = c.get("bpp", 0.5)
    line = {
        "impl": "reference", "metric": "encode megapixels/sec at %.3g bpp" % bpp, "value": round(value, 3),
        "unit": "MP/s", "n_gpus": args.gpus, "steps": len(vals), "warmup": args.warmup,
        "ms_per_step": round(pixels / 1e6 / value * 1e3, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic" if not c.get("image") else "reference fixture image",
        "config": {"workload": workload_name(c), "sample_images_per_step": n_images},
        "decode": {"value": round(statistics.mean(dvals), 3), "unit": "MP/s"},
        "cpu_baseline": {"value": round(value, 3), "unit": "MP/s", "cores": used, "kind": "port",
                         "sample": f"{n_images} images of the workload per step, one image per process; " + CPU_SAMPLE_NOTE,
                         "detail": detail},
        "e2e": {"value": round(value, 3), "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU arm
def bind_to_gpu_numa_node(local_rank):
    """Pin this rank's threads (and, by first touch, its pinned host buffers) to the NUMA node of its GPU: with 8
    ranks pushing ~3 GB per step each, host memory bandwidth and the PCIe root are what the end-to-end number hits.
    Returns a short description for the JSON line."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_rank)
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open(path).read().strip())
        if node < 0:
            return {"numa_node": None, "note": "no NUMA affinity reported for the GPU"}
        cpus = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
        ids = []
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.extend(range(int(a), int(b or a) + 1))
        allowed = sorted(set(ids) & os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
        return {"numa_node": node, "cpus": len(allowed)}
    except Exception as e:   # not fatal: the benchmark runs unbound
        return {"numa_node": None, "note": f"not bound ({type(e).__name__})"}


def load_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback"


def make_pixels(c, rank, dev):
    import torch
    from spiht_b200.utils import imload, synthetic_images
    if c.get("image"):
        import numpy as np
        im = imload(os.path.join(ROOT, c["image"]))
        return torch.from_numpy(np.ascontiguousarray(im))[None].expand(c["batch"], -1, -1, -1).contiguous().to(dev)
    chunk = 16 if c["h"] <= 2048 else 1
    return synthetic_images(c["batch"], 3, c["h"], c["w"], seed=1000 * c["config"] + rank, device=dev, chunk=chunk)


def run_b200(args):
    import torch
    import spiht_b200 as spiht
    from spiht_b200 import _lib, batch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the SPIHT hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(local_rank)
    use_dist = world > 1
    dist = None
    if use_dist:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    c = resolve(args)
    if "mixed" in c:
        return run_b200_mixed(args, c, rank, world, local_rank, dev, dist, numa)
    from spiht_b200 import dist as sdist

    B, H, W = c["batch"], c["h"], c["w"]
    settings = spiht.SpihtSettings(wavelet=c["wavelet"], mode=c["mode"], **c["settings"])
    g = _lib.plan(H, W, c["wavelet"], c["mode"], None)
    max_bits = c["max_bits"]
    C = 3
    pixels = make_pixels(c, rank, dev)
    stride = batch.stream_stride(max_bits, C, g)
    coeffs = torch.empty((B, C, g.enc_h, g.enc_w), dtype=torch.int32, device=dev)
    streams = torch.zeros((B, stride), dtype=torch.uint8, device=dev)
    ctx = _lib.get_context(local_rank)

    def step():
        out = batch.encode_images(pixels, g, settings, max_bits, out_stride=stride, coeffs=coeffs, out=streams)
        if use_dist:
            sdist.gather_lengths(out[1], out[2])
        return out

    def barrier():
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    # Warm-up: exactly max(W, 3) steps on every rank (every step posts a collective when N > 1, so the count
    # must never depend on a per-rank clock).  nvidia-smi needs ~0.8 s under load to return a few samples and
    # the timed region alone (K steps of a few ms) is shorter than one sampling period, so the warm-up is
    # followed by untimed "soak" steps whose number rank 0 decides from its own clock and broadcasts.
    clocks = ClockSampler(local_rank)
    clocks.start()
    n_warm = max(args.warmup, 3)
    t_warm = time.perf_counter()
    for _ in range(n_warm):
        res = step()
        torch.cuda.synchronize()
    per_step = (time.perf_counter() - t_warm) / n_warm
    n_soak = torch.tensor([min(400, max(0, int(0.8 / max(per_step, 1e-4)) - n_warm))], dtype=torch.int64, device=dev)
    if use_dist:
        dist.broadcast(n_soak, src=0)
    n_soak = int(n_soak.item())
    for _ in range(n_soak):
        res = step()
    barrier()

    # ---- timed region: K steps, CUDA events on the launching stream, stage timers on
    ctx.profile(True)
    ctx.profile_read(reset=True)
    launches0 = ctx.launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        res = step()
    e1.record()
    barrier()
    enc_ms = e0.elapsed_time(e1)
    clock_info = clocks.stop()
    launches = ctx.launch_count() - launches0
    stages = ctx.profile_read(reset=True)
    ctx.profile(False)
    nbits, max_n = res[1], res[2]
    assert int(nbits.min().item()) > 0

    # ---- decode (same metric, mirror path)
    dec_ms = None
    psnr = None
    if not args.no_decode:
        nbytes = (nbits + 7) // 8
        dec_pix = torch.empty((B, C, g.rec_h, g.rec_w), dtype=torch.float32, device=dev)

        def dstep():
            # pixels are the product: the coefficient array is scratch (SPIHTB_OPT_SCRATCH_COEFFS), as in
            # spiht_b200.decode_images
            return batch.decode_images(streams, nbytes, max_n, C, g, settings, dtype=torch.float32, coeffs=coeffs,
                                       out=dec_pix, scratch_coeffs=True)
        for _ in range(max(args.warmup, 3)):
            dstep()
        barrier()
        ctx.profile(True)
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d0.record()
        for _ in range(args.steps):
            dstep()
        d1.record()
        barrier()
        dec_ms = d0.elapsed_time(d1)
        dstages = ctx.profile_read(reset=True)
        ctx.profile(False)
        stages.update({k: v for k, v in dstages.items() if v[1]})
        sq = 0.0
        nb = max(1, min(B, (1 << 28) // (C * H * W)))   # in slices: no second batch-sized temporary
        for lo in range(0, B, nb):
            sq += float(((dec_pix[lo:lo + nb, :, :H, :W] - pixels[lo:lo + nb].float()) ** 2).sum().item())
        psnr = float(10 * torch.log10(torch.tensor(1.0 / (sq / (B * C * H * W)))).item())
        del dec_pix

    # ---- end to end through the public API: pinned host pixels in, host bytes out
    e2e = None
    e2e_u8_s = 0.0
    if not args.no_e2e:
        e2e_B = min(B, 256) if H * W <= 2048 * 2048 else min(B, 4)
        host_pixels = torch.empty((e2e_B, C, H, W), dtype=torch.float32).pin_memory()
        host_pixels.copy_(pixels[:e2e_B])
        torch.cuda.synchronize()
        # the bare host->device copy of the same buffer, per rank (what the PCIe root / host memory gives this rank
        # while every other rank does the same)
        devbuf = torch.empty(host_pixels.shape, dtype=torch.float32, device=dev)
        barrier()
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0.record()
        devbuf.copy_(host_pixels, non_blocking=True)
        h1.record()
        torch.cuda.synchronize()
        h2d_gbs = host_pixels.numel() * 4 / (h0.elapsed_time(h1) / 1e3) / 1e9
        del devbuf
        spiht.encode_images(host_pixels, settings, None, max_bits)      # warm
        barrier()
        t0 = time.perf_counter()
        d2h = 0
        for _ in range(args.e2e_steps):
            encs = spiht.encode_images(host_pixels, settings, None, max_bits)
            d2h = sum(len(e.encoded_bytes) for e in encs) + 12 * len(encs)
        torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - t0) / args.e2e_steps
        # the same call with the pixels as image bytes (uint8, what an image file holds): the library applies the
        # reference loader's 1/255 scaling on the device, and the host->device copy is a quarter of the float32 one
        host_u8 = (pixels[:e2e_B].float() * 255.0).round().clamp_(0, 255).to(torch.uint8).cpu().pin_memory()
        torch.cuda.synchronize()
        spiht.encode_images(host_u8, settings, None, max_bits)           # warm
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            spiht.encode_images(host_u8, settings, None, max_bits)
        torch.cuda.synchronize()
        e2e_u8_s = (time.perf_counter() - t0) / args.e2e_steps
        e2e = (e2e_s, e2e_B, d2h, h2d_gbs)
    red = torch.tensor([e2e[0] if e2e else 0.0, enc_ms, dec_ms or 0.0, e2e_u8_s], dtype=torch.float64, device=dev)
    h2d_all = torch.tensor([e2e[3] if e2e else 0.0], dtype=torch.float64, device=dev)
    if use_dist:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        gathered = [torch.zeros_like(h2d_all) for _ in range(world)]
        dist.all_gather(gathered, h2d_all)
        h2d_ranks = [round(float(t.item()), 1) for t in gathered]
    else:
        h2d_ranks = [round(float(h2d_all.item()), 1)]
    e2e_s, enc_ms, dec_ms_max, e2e_u8_s = [float(v) for v in red.tolist()]

    if rank != 0:
        if use_dist:
            dist.destroy_process_group()
        return

    mp_step = world * B * H * W / 1e6
    ms_per_step = enc_ms / args.steps
    value = mp_step / (ms_per_step / 1e3)

    # ---- roofline of the dominant kernel (algorithmic bytes)
    peak, peak_src = load_peak()
    es = pixels.element_size()
    px_bytes = es * C * H * W
    coef_bytes = 4 * C * g.enc_h * g.enc_w
    det1 = 4 * C * (g.enc_h * g.enc_w - g.off_h[0] * g.off_w[0])      # level-1 detail blocks (incl. gaps)
    fused12 = ctx.forward_path() == 12                                 # levels 1+2 in one kernel (csrc/dwt_fwd2.cu)
    if fused12 and g.levels > 1:
        det1 += 4 * C * (g.off_h[0] * g.off_w[0] - g.off_h[1] * g.off_w[1])   # + level-2 detail blocks
    stream_bytes = (max_bits + 7) // 8
    alg = {  # bytes per image and launch
        "dwt_fwd_level1": px_bytes + det1,
        "dwt_fwd_rest": coef_bytes - det1,
        # the base pass of the pyramid (the one algorithmic read of the coefficient array) is
        # fused into the forward transform's epilogue; what is left under this timer is the fix-up of cells that
        # straddle two warps' tiles (implementation traffic)
        "pyramid_base": 0,
        "pyramid_rest": 0,
        "spiht_encode": stream_bytes,
        "spiht_decode": coef_bytes + stream_bytes,
        "dwt_inv_coarse": coef_bytes - det1,
        "dwt_inv_level1": 4 * C * H * W + det1,
    }
    # a stage may be launched several times per step (image groups): per-step time = total / steps
    stage_ms = {k: (v[0] / args.steps if v[1] else 0.0) for k, v in stages.items()}
    enc_stages = ("dwt_fwd_level1", "dwt_fwd_rest", "pyramid_base", "pyramid_rest", "spiht_encode")
    dom = max(enc_stages, key=lambda k: stage_ms[k])
    dom_gbs = alg[dom] * B / (stage_ms[dom] / 1e3) / 1e9 if stage_ms[dom] > 0 else 0.0
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get(f"config{c['config']}", {}).get(dom)
    except Exception:
        pass
    A = px_bytes + 2 * coef_bytes + stream_bytes
    step_gbs = A * B / (ms_per_step / 1e3) / 1e9

    line = {
        "metric": "encode megapixels/sec at %.3g bpp" % c["bpp"], "value": round(value, 1), "unit": "MP/s",
        "n_gpus": world, "steps": args.steps, "warmup": n_warm, "soak_steps": n_soak,
        "ms_per_step": round(ms_per_step, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64/int32",
        "data": "synthetic" if not c.get("image") else "reference fixture image (replicated over the batch)",
        "config": {"workload": workload_name(c), "global_batch": world * B, "max_bits": max_bits,
                   "quantization_scale": settings.quantization_scale, "levels": g.levels,
                   "coeff_array": [C, g.enc_h, g.enc_w],
                   "l2": ("inputs exceed L2 (%.2f GB pixels + %.2f GB coefficients per step)"
                          % (px_bytes * B / 1e9, coef_bytes * B / 1e9)) if (px_bytes + coef_bytes) * B > 200e6 else
                         "working set fits L2: one small image per step (launch-latency bound by nature); no flush",
                   "parallelism": f"images sharded over {world} GPU(s), no data-path collective"},
        "clocks": clock_info,
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": round(dom_gbs, 1), "peak": peak, "unit": "GB/s",
                     "frac": round(dom_gbs / peak, 4), "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_image": alg[dom], "ms_per_launch": round(stage_ms[dom], 4),
                     "kernel_name": ("dwt_fwd12_kernel (levels 1+2 fused, TMA-staged)" if fused12 and dom == "dwt_fwd_level1"
                                     else dom)},
        "step_roofline": {"achieved": round(step_gbs, 1), "peak": peak, "unit": "GB/s",
                          "frac": round(step_gbs / peak, 4), "algorithmic_bytes_per_image": A,
                          "strict_io_frac": round((px_bytes + stream_bytes) * B / (ms_per_step / 1e3) / 1e9 / peak, 4)},
        "stages_ms": {k: round(v, 4) for k, v in stage_ms.items() if v},
        "stages_gbs": {k: round(alg[k] * B / (v / 1e3) / 1e9, 1) for k, v in stage_ms.items() if v and alg[k]},
    }
    if e2e:
        _, e2e_B, d2h, _ = e2e
        line["e2e"] = {"value": round(world * e2e_B * H * W / 1e6 / e2e_s, 1), "unit": "MP/s",
                       "h2d_bytes_per_step": 4 * C * H * W * e2e_B, "d2h_bytes_per_step": d2h,
                       "images_per_step": e2e_B,
                       "api": "spiht_b200.encode_images(pinned host float32 [B,3,H,W]) -> list[EncodingResult]",
                       # what bounds it: the bare pinned->device copy of the same buffer on every rank at once
                       "h2d_gbs_per_rank": h2d_ranks,
                       "h2d_copy_share_of_step": round(4 * C * H * W * e2e_B / (min(h2d_ranks) * 1e9) / e2e_s, 3)
                       if min(h2d_ranks) > 0 else None,
                       "numa": numa}
        line["e2e_uint8"] = {"value": round(world * e2e_B * H * W / 1e6 / e2e_u8_s, 1), "unit": "MP/s",
                             "h2d_bytes_per_step": C * H * W * e2e_B,
                             "note": "same call, pixels as uint8 image bytes (round(255 x)); not the headline: other input values"}
    else:
        line["e2e"] = None
    if dec_ms:
        dstep = dec_ms_max / args.steps
        line["decode"] = {"value": round(mp_step / (dstep / 1e3), 1), "unit": "MP/s", "ms_per_step": round(dstep, 4),
                          "psnr_db": round(psnr, 2),
                          "step_roofline_frac": round(A * B / (dstep / 1e3) / 1e9 / peak, 4)}
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n_images = cpu_sample_size(c, cores, args.cpu_images)
        enc_mps, dec_mps, detail = cpu_reference_sample(c, n_images, cores)
        line["cpu_baseline"] = {"value": round(enc_mps, 3), "unit": "MP/s", "cores": min(cores, n_images), "kind": "port",
                                "decode_value": round(dec_mps, 3),
                                "sample": f"{n_images} images of the workload, one image per process; " + CPU_SAMPLE_NOTE,
                                "detail": detail}
    print(json.dumps(line), flush=True)
    if use_dist:
        dist.destroy_process_group()


def run_b200_mixed(args, c, rank, world, local_rank, dev, dist, numa):
    """configs[4]: one mixed-size batch per GPU (device-resident float32), every (mode, bpp) of the sweep encoded and
    decoded; images of one size go through one library call (mixed shapes are grouped, as spiht.encode_images does)."""
    import torch
    import spiht_b200 as spiht
    from spiht_b200 import _lib, batch
    from spiht_b200.utils import synthetic_images
    use_dist = world > 1
    ctx = _lib.get_context(local_rank)
    groups = []
    for s, n in c["mixed"]:
        groups.append((s, n, synthetic_images(n, 3, s, s, seed=5000 + 10 * rank + s % 97, device=dev,
                                              chunk=16 if s <= 2048 else 1)))
    pix_total = sum(n * s * s for s, n, _ in groups)
    peak, peak_src = load_peak()

    def barrier():
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local_rank)
    clocks.start()
    sweep = []
    launches = 0
    tot_enc_ms = tot_dec_ms = 0.0
    alg_total = 0
    for mode in c["modes"]:
        st = spiht.SpihtSettings(wavelet=c["wavelet"], mode=mode)
        plans = [(s, n, px, _lib.plan(s, s, c["wavelet"], mode, None)) for s, n, px in groups]
        for bpp in c["bpps"]:
            bufs = []
            for s, n, px, g in plans:
                mb = int(s * s * bpp)
                stride = batch.stream_stride(mb, 3, g)
                bufs.append((mb, stride, torch.empty((n, 3, g.enc_h, g.enc_w), dtype=torch.int32, device=dev),
                             torch.zeros((n, stride), dtype=torch.uint8, device=dev),
                             torch.empty((n, 3, g.rec_h, g.rec_w), dtype=torch.float32, device=dev)))

            def enc_step():
                outs = []
                for (s, n, px, g), (mb, stride, co, out, _) in zip(plans, bufs):
                    outs.append(batch.encode_images(px, g, st, mb, out_stride=stride, coeffs=co, out=out))
                return outs

            def dec_step(outs):
                for (s, n, px, g), (mb, stride, co, out, rec), o in zip(plans, bufs, outs):
                    batch.decode_images(o[0], (o[1] + 7) // 8, o[2], 3, g, st, dtype=torch.float32, coeffs=co, out=rec,
                                        scratch_coeffs=True)

            for _ in range(max(args.warmup, 3)):
                outs = enc_step()
                dec_step(outs)
            barrier()
            l0 = ctx.launch_count()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            for _ in range(args.steps):
                outs = enc_step()
            e1.record()
            for _ in range(args.steps):
                dec_step(outs)
            e2.record()
            barrier()
            launches += ctx.launch_count() - l0
            t = torch.tensor([e0.elapsed_time(e1), e1.elapsed_time(e2)], dtype=torch.float64, device=dev)
            if use_dist:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            enc_ms, dec_ms = (float(v) / args.steps for v in t.tolist())
            A = sum(n * (4 * 3 * s * s + 2 * 4 * 3 * g.enc_h * g.enc_w + (int(s * s * bpp) + 7) // 8)
                    for s, n, px, g in plans)
            psnr = min(float(10 * torch.log10(1.0 / torch.mean((b[4][:, :, :s, :s] - px) ** 2)).item())
                       for (s, n, px, g), b in zip(plans, bufs))
            sweep.append({"mode": mode, "bpp": bpp, "encode_mps": round(world * pix_total / 1e6 / (enc_ms / 1e3), 1),
                          "decode_mps": round(world * pix_total / 1e6 / (dec_ms / 1e3), 1),
                          "encode_ms": round(enc_ms, 4), "decode_ms": round(dec_ms, 4),
                          "encode_roofline_frac": round(A / (enc_ms / 1e3) / 1e9 / peak, 4),
                          "decode_roofline_frac": round(A / (dec_ms / 1e3) / 1e9 / peak, 4),
                          "min_psnr_db": round(psnr, 2)})
            tot_enc_ms += enc_ms
            tot_dec_ms += dec_ms
            alg_total += A
            del bufs
    clock_info = clocks.stop()
    if rank != 0:
        if use_dist:
            dist.destroy_process_group()
        return
    ncomb = len(sweep)
    value = world * pix_total * ncomb / 1e6 / (tot_enc_ms / 1e3)
    frac = alg_total / (tot_enc_ms / 1e3) / 1e9 / peak
    line = {
        "metric": "encode megapixels/sec over the bpp / mode sweep", "value": round(value, 1), "unit": "MP/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(tot_enc_ms / ncomb, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64/int32", "data": "synthetic",
        "config": {"workload": workload_name(c), "images_per_gpu": sum(n for _, n, _ in groups),
                   "megapixels_per_gpu": round(pix_total / 1e6, 1),
                   "l2": "inputs exceed L2 (%.2f GB pixels per step)" % (pix_total * 12 / 1e9),
                   "parallelism": f"images sharded over {world} GPU(s), no data-path collective"},
        "clocks": clock_info, "gpu_launches": launches, "e2e": None,
        "decode": {"value": round(world * pix_total * ncomb / 1e6 / (tot_dec_ms / 1e3), 1), "unit": "MP/s",
                   "ms_per_step": round(tot_dec_ms / ncomb, 4)},
        "step_roofline": {"achieved": round(frac * peak, 1), "peak": peak, "unit": "GB/s", "frac": round(frac, 4),
                          "peak_source": peak_src},
        "roofline": {"bound": "hbm", "kernel": "whole step (mixed shapes: one launch sequence per size)",
                     "achieved": round(frac * peak, 1), "peak": peak, "unit": "GB/s", "frac": round(frac, 4),
                     "traffic": None, "peak_source": peak_src},
        "sweep": sweep,
    }
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        enc_mps, dec_mps, _, n_images, used = _cpu_mixed(c, cores)
        line["cpu_baseline"] = {"value": round(enc_mps, 3), "unit": "MP/s", "cores": used, "kind": "port",
                                "decode_value": round(dec_mps, 3),
                                "sample": f"{n_images} images: one of every size and mode at 0.5 bpp, timed one at a time "
                                          "and scaled to one image per process; " + CPU_SAMPLE_NOTE}
    print(json.dumps(line), flush=True)
    if use_dist:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
