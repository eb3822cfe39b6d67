#!/usr/bin/env python
"""bench.py -- encode / decode megapixels per second of the SPIHT hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on the host cores

Workload at N=1: BASELINE.json configs[1] -- a batch of 256 synthetic 1024x1024 RGB images, bior2.2,
mode=reflect, 0.5 bpp (max_bits = 524288 per image), quantisation scale 50.  A "step" is one pass of
the hot path (RGB -> 7-level DWT -> int32 quantise -> pyramid -> SPIHT encode) over that batch.  With
N > 1 every rank owns its own batch of the same size (images are independent: weak scaling, no
collective on the data path; an NCCL all-gather of the per-image stream lengths closes every step).

One JSON line is printed by rank 0 (see the task contract): `value` is whole-job encode MP/s with the
inputs resident in HBM, `e2e` the same metric through the public API with pinned HOST buffers in and
host bytes out, `roofline` the achieved HBM GB/s of the dominant kernel against MEASURED_PEAKS.json,
`cpu_baseline` the CPU oracle (a port of the reference algorithm) timed on the host cores.
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="images per GPU")
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--bpp", type=float, default=0.5)
    ap.add_argument("--wavelet", default="bior2.2")
    ap.add_argument("--mode", default="reflect")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-images", type=int, default=0, help="images in the CPU baseline sample (0 = one per core)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-decode", action="store_true")
    return ap.parse_args()


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
def _cpu_worker(args):
    idx, size, bpp, wavelet, mode, seed = args
    import numpy as np
    sys.path.insert(0, ROOT)
    from oracle import spiht_oracle, wrapper_ref
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    rng = np.random.default_rng(seed + idx)
    fy = np.fft.fftfreq(size)[:, None]
    fx = np.fft.fftfreq(size)[None, :]
    f = np.sqrt(fy * fy + fx * fx)
    f[0, 0] = 1.0
    fields = np.fft.ifft2(np.fft.fft2(rng.normal(size=(4, size, size))) / f).real
    img = 0.8 * fields[:1] + 0.2 * fields[1:]
    img = (img - img.min(axis=(1, 2), keepdims=True)) / np.ptp(img, axis=(1, 2), keepdims=True)
    img = img.astype(np.float32).astype(np.float64)
    max_bits = int(size * size * bpp)
    t0 = time.perf_counter()
    arr, ll_h, ll_w = wrapper_ref.forward_coeffs(img, wavelet, mode, fast=True)   # compiled transform (oracle/dwt_fast.c)
    t1 = time.perf_counter()
    data, max_n = spiht_oracle.encode(arr, ll_h, ll_w, max_bits)
    t2 = time.perf_counter()
    rec = spiht_oracle.decode(data, max_n, 3, arr.shape[1], arr.shape[2], ll_h, ll_w)
    t3 = time.perf_counter()
    wrapper_ref.inverse_coeffs(rec, size, size, wavelet, mode, fast=True)
    t4 = time.perf_counter()
    return (t1 - t0, t2 - t1, t3 - t2, t4 - t3)


def cpu_reference_sample(size, bpp, wavelet, mode, n_images, cores, seed=4242):
    """the oracle (C restatements of the Rust coder and of the PyWavelets transform, float64) on n_images images, one image per
    process over `cores` processes.  Returns (encode MP/s, decode MP/s, detail dict)."""
    from oracle import dwt_fast, spiht_oracle
    spiht_oracle.build()
    dwt_fast.lib()
    jobs = [(i, size, bpp, wavelet, mode, seed) for i in range(n_images)]
    with mp.get_context("spawn").Pool(cores) as pool:
        pool.map(_cpu_worker, [(0, 64, bpp, wavelet, mode, seed)] * cores)   # warm the workers (imports, build)
        t0 = time.perf_counter()
        res = pool.map(_cpu_worker, jobs, chunksize=1)
        wall = time.perf_counter() - t0
    mp_total = n_images * size * size / 1e6
    enc = sum(r[0] + r[1] for r in res)
    dec = sum(r[2] + r[3] for r in res)
    # wall includes image synthesis; throughput is over the codec time only, scaled to the cores in use
    enc_mps = mp_total / (enc / min(cores, n_images))
    dec_mps = mp_total / (dec / min(cores, n_images))
    detail = {"images": n_images, "wall_s": round(wall, 2),
              "per_image_s": {"dwt_quant": round(statistics.mean(r[0] for r in res), 3),
                              "spiht_encode": round(statistics.mean(r[1] for r in res), 3),
                              "spiht_decode": round(statistics.mean(r[2] for r in res), 3),
                              "inverse_dwt": round(statistics.mean(r[3] for r in res), 3)}}
    return enc_mps, dec_mps, detail


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_images = args.cpu_images or 4 * cores
    workload = workload_name(args)
    vals, dvals = [], []
    t_all = time.perf_counter()
    detail = None
    for step in range(args.warmup + args.steps):
        enc_mps, dec_mps, detail = cpu_reference_sample(args.size, args.bpp, args.wavelet, args.mode, n_images, cores,
                                                        seed=4242 + 100 * step)
        if step >= args.warmup:
            vals.append(enc_mps)
            dvals.append(dec_mps)
        if time.perf_counter() - t_all > 240 and len(vals) >= 1:
            break
    value = statistics.mean(vals)
    pixels = n_images * args.size * args.size
    line = {
        "impl": "reference", "metric": "encode megapixels/sec at %.3g bpp" % args.bpp, "value": round(value, 3),
        "unit": "MP/s", "n_gpus": args.gpus, "steps": len(vals), "warmup": args.warmup,
        "ms_per_step": round(pixels / 1e6 / value * 1e3, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "sample_images_per_step": n_images},
        "decode": {"value": round(statistics.mean(dvals), 3), "unit": "MP/s"},
        "cpu_baseline": {"value": round(value, 3), "unit": "MP/s", "cores": cores, "kind": "port",
                         "sample": f"{n_images} images of the workload per step, one image per process; "
                                   "oracle/spiht_ref.c (restated Rust coder) + oracle/dwt_fast.c (restated PyWavelets transform, float64)",
                         "detail": detail},
        "e2e": {"value": round(value, 3), "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_name(args):
    return (f"batch of {args.batch} synthetic {args.size}x{args.size} RGB images per GPU, {args.wavelet} "
            f"{args.mode}, {args.bpp:g} bpp (BASELINE.json configs[1])")


# ----------------------------------------------------------------------------- GPU arm
def run_b200(args):
    import torch
    import spiht_b200 as spiht
    from spiht_b200 import _lib, batch
    from spiht_b200.utils import synthetic_images

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the SPIHT hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    use_dist = world > 1
    if use_dist:
        import torch.distributed as dist
        from spiht_b200 import dist as sdist
        dist.init_process_group("nccl", device_id=dev)

    B, S = args.batch, args.size
    settings = spiht.SpihtSettings(wavelet=args.wavelet, mode=args.mode)
    g = _lib.plan(S, S, args.wavelet, args.mode, None)
    max_bits = int(S * S * args.bpp)
    C = 3
    pixels = synthetic_images(B, C, S, S, seed=1000 * 2 + rank, device=dev)
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
    if not args.no_decode:
        nbytes = (nbits + 7) // 8
        dec_pix = torch.empty((B, C, g.rec_h, g.rec_w), dtype=torch.float32, device=dev)

        def dstep():
            return batch.decode_images(streams, nbytes, max_n, C, g, settings, dtype=torch.float32, coeffs=coeffs,
                                       out=dec_pix)
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
        psnr = float(10 * torch.log10(1.0 / torch.mean((dec_pix[:, :, :S, :S] - pixels) ** 2)).item())
    else:
        psnr = None

    # ---- end to end through the public API: pinned host pixels in, host bytes out
    e2e_B = B
    host_pixels = torch.empty((e2e_B, C, S, S), dtype=torch.float32).pin_memory()
    host_pixels.copy_(pixels[:e2e_B])
    torch.cuda.synchronize()
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
    host_u8 = (pixels[:e2e_B] * 255.0).round().clamp_(0, 255).to(torch.uint8).cpu().pin_memory()
    torch.cuda.synchronize()
    spiht.encode_images(host_u8, settings, None, max_bits)           # warm
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        spiht.encode_images(host_u8, settings, None, max_bits)
    torch.cuda.synchronize()
    e2e_u8_s = (time.perf_counter() - t0) / args.e2e_steps
    e2e_t = torch.tensor([e2e_s, enc_ms, dec_ms or 0.0, e2e_u8_s], dtype=torch.float64, device=dev)
    if use_dist:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_s, enc_ms, dec_ms_max, e2e_u8_s = [float(v) for v in e2e_t.tolist()]

    if rank != 0:
        if use_dist:
            dist.destroy_process_group()
        return

    mp_step = world * B * S * S / 1e6
    ms_per_step = enc_ms / args.steps
    value = mp_step / (ms_per_step / 1e3)

    # ---- roofline of the dominant kernel (algorithmic bytes)
    peak, peak_src = FALLBACK_HBM_GBS, "fallback"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak = float(json.load(f)["hbm_gbs"])
            peak_src = "measured"
    except Exception:
        pass
    px_bytes = 4 * C * S * S
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
        # fused into the forward transform's epilogue; what is left under this timer is a zero fill of the
        # byte planes and the fix-up of cells that straddle two warps' tiles (implementation traffic)
        "pyramid_base": 0,
        "pyramid_rest": 0,
        "spiht_encode": stream_bytes,
        "spiht_decode": coef_bytes + stream_bytes,
        "dwt_inv_coarse": coef_bytes - det1,
        "dwt_inv_level1": px_bytes + det1,
    }
    # a stage may be launched several times per step (image groups): per-step time = total / steps
    stage_ms = {k: (v[0] / args.steps if v[1] else 0.0) for k, v in stages.items()}
    enc_stages = ("dwt_fwd_level1", "dwt_fwd_rest", "pyramid_base", "pyramid_rest", "spiht_encode")
    dom = max(enc_stages, key=lambda k: stage_ms[k])
    dom_gbs = alg[dom] * B / (stage_ms[dom] / 1e3) / 1e9 if stage_ms[dom] > 0 else 0.0
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get(dom)
    except Exception:
        pass
    A = px_bytes + 2 * coef_bytes + stream_bytes
    step_gbs = A * B / (ms_per_step / 1e3) / 1e9

    line = {
        "metric": "encode megapixels/sec at %.3g bpp" % args.bpp, "value": round(value, 1), "unit": "MP/s",
        "n_gpus": world, "steps": args.steps, "warmup": n_warm, "soak_steps": n_soak, "ms_per_step": round(ms_per_step, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64/int32",
        "data": "synthetic",
        "config": {"workload": workload_name(args), "global_batch": world * B, "max_bits": max_bits,
                   "quantization_scale": 50.0, "levels": g.levels, "coeff_array": [C, g.enc_h, g.enc_w],
                   "l2": "inputs exceed L2 (%.1f GB pixels + %.1f GB coefficients per step)"
                         % (px_bytes * B / 1e9, coef_bytes * B / 1e9),
                   "parallelism": f"images sharded over {world} GPU(s), no data-path collective"},
        "clocks": clock_info,
        "e2e": {"value": round(world * e2e_B * S * S / 1e6 / e2e_s, 1), "unit": "MP/s",
                "h2d_bytes_per_step": px_bytes * e2e_B, "d2h_bytes_per_step": d2h,
                "api": "spiht_b200.encode_images(pinned host float32 [B,3,H,W]) -> list[EncodingResult]"},
        "e2e_uint8": {"value": round(world * e2e_B * S * S / 1e6 / e2e_u8_s, 1), "unit": "MP/s",
                      "h2d_bytes_per_step": C * S * S * e2e_B,
                      "note": "same call, pixels as uint8 image bytes (round(255 x)); not the headline: other input values"},
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
    if dec_ms:
        line["decode"] = {"value": round(mp_step / (dec_ms_max / args.steps / 1e3), 1), "unit": "MP/s",
                          "ms_per_step": round(dec_ms_max / args.steps, 4), "psnr_db": round(psnr, 2)}
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n_images = args.cpu_images or 4 * cores
        enc_mps, dec_mps, detail = cpu_reference_sample(S, args.bpp, args.wavelet, args.mode, n_images, cores)
        line["cpu_baseline"] = {"value": round(enc_mps, 3), "unit": "MP/s", "cores": cores, "kind": "port",
                                "decode_value": round(dec_mps, 3),
                                "sample": f"{n_images} images of the workload, one image per process; "
                                          "oracle/spiht_ref.c (restated Rust coder) + oracle/dwt_fast.c (restated PyWavelets transform, float64)",
                                "detail": detail}
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
