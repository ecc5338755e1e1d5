#!/usr/bin/env python
"""Benchmark of the deep-spectral hot path: images/sec for extract_features + extract_eigs
(480x480, dino_vits16, K=5, dense affinity -- BASELINE.json configs[1]) on N B200s, one process per GPU.

    python bench.py --gpus 1 --steps 4 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's CPU path (oracle port) on the host cores

A step = one pass of the hot path over one batch of `--images-per-step` synthetic images per GPU (weak scaling).
Rank 0 prints ONE JSON line. `value` is timed with the uint8 images already resident in HBM; `e2e` goes through
the public host-buffer call (pinned host images in, eigenvectors out, copies inside the timed region)."""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
PKG = "deep-spectral-segmentation_b200"

METRIC = "images/sec (features+eigs, 480px dino_vits16 K=5)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--images-per-step", type=int, default=256)
    ap.add_argument("--size", type=int, default=480)
    ap.add_argument("--K", type=int, default=5)
    ap.add_argument("--model", default="dino_vits16")
    ap.add_argument("--vit-batch", type=int, default=0, help="images per ViT launch sequence (0 = the whole step)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="images of the CPU-baseline sample (0 = 2 x workers)")
    ap.add_argument("--parity-sample", type=int, default=4)
    ap.add_argument("--ref-images-per-step", type=int, default=0, help="0 = 2 x worker processes")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed regions. The sampler process is started before
    the warm-up (it needs a few hundred ms to come up); only samples whose timestamp falls inside a timed window
    (mark_begin / mark_end) are kept."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.path, self.windows, self._t0 = index, None, None, [], None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def mark_begin(self):
        self._t0 = time.time()

    def mark_end(self):
        if self._t0 is not None:
            self.windows.append((self._t0, time.time()))
            self._t0 = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        import datetime
        rows = []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in Path(self.path).read_text().splitlines():
                p = [x.strip() for x in line.split(",")]
                if len(p) < 8:
                    continue
                try:
                    ts = datetime.datetime.strptime(p[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                except Exception:
                    ts = None
                rows.append((ts, float(p[1]), float(p[2]), [n for n, v in zip(names, p[4:8]) if v.lower().startswith("active")]))
            os.unlink(self.path)
        except Exception:
            pass
        inside = [r for r in rows if r[0] is not None and any(a - 0.02 <= r[0] <= b + 0.02 for a, b in self.windows)]
        used, where = (inside, "timed regions") if inside else (rows, "whole run (no sample fell inside the timed regions)")
        sm = sorted(r[1] for r in used)
        reasons = sorted({n for r in used for n in r[3]})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": used[-1][2] if used else None,
                "reasons": reasons, "samples": len(used), "window": where}


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.is_file():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "tflops_burst": d["bf16_tflops"], "tflops_sustained": d["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


# ---------------------------------------------------------------------------------------------------------------
def cpu_reference_images_per_sec(images_u8, model_name, K, sd, want_outputs=False):
    """The reference's path on the host CPU (oracle port): fp32 eager DINO ViT + the reference's scipy eigsh route.
    Returns (images/s, per-image outputs)."""
    import torch
    from oracle import dino_vit, eigs_ref
    torch.set_grad_enabled(False)
    ref = dino_vit.DinoViT(dino_vit.cfg_for(model_name)).eval()
    ref.load_state_dict({k: v.float().cpu() for k, v in sd.items()})
    P = ref.cfg.patch
    outs = []
    t0 = time.perf_counter()
    t_vit = 0.0
    for img in images_u8:
        a = time.perf_counter()
        k = ref.forward_k(dino_vit.preprocess_u8(img, P))
        t_vit += time.perf_counter() - a
        ev, vec = eigs_ref.extract_eig(k[0], K)
        if want_outputs:
            outs.append((k[0], ev, vec))
    dt = time.perf_counter() - t0
    return len(images_u8) / dt, outs, {"vit_s_per_image": t_vit / len(images_u8), "eigs_s_per_image": (dt - t_vit) / len(images_u8)}


# ---- multi-process CPU arm: the reference's own parallel knob is a process pool over images
# (extract_utils.py:138-148, `--multiprocessing N`); each worker runs the fp32 ViT + scipy eigsh route with a few
# BLAS threads so that all host cores are busy.
_W = {}


def _ref_worker_init(model_name, threads):
    import torch
    torch.set_num_threads(threads)
    torch.set_grad_enabled(False)
    from oracle import dino_vit
    vit = importlib.import_module(PKG + ".vit")
    m = dino_vit.DinoViT(dino_vit.cfg_for(model_name)).eval()
    m.load_state_dict(vit.random_state_dict(model_name, 0))
    _W["model"] = m


def _ref_worker_task(job):
    seed, size, K = job
    from oracle import dino_vit, eigs_ref
    synth = importlib.import_module(PKG + ".synth")
    m = _W["model"]
    img = synth.blobs_image(size, size, seed)
    t0 = time.perf_counter()
    k = m.forward_k(dino_vit.preprocess_u8(img, m.cfg.patch))
    t1 = time.perf_counter()
    eigs_ref.extract_eig(k[0], K)
    return (t1 - t0, time.perf_counter() - t1)


def cpu_pool_images_per_sec(model_name, size, K, n_images, steps, warmup, threads_per_worker=2):
    """images/s of the CPU path with every host core busy: cores/threads_per_worker worker processes (2 threads per
    worker measured best on the 2 x 32-core / 128-thread host of the B200 box: 9.3 img/s vs 7.6 at 4, 5.9 at 8)."""
    import multiprocessing as mp
    from concurrent.futures import ProcessPoolExecutor
    cores = os.cpu_count() or 1
    workers = max(1, min(cores // threads_per_worker, n_images, 64))
    ctx = mp.get_context("spawn")
    # numpy/scipy's BLAS (OpenBLAS) sizes its own thread pool from the environment at import time: without this
    # every worker would start one BLAS thread per host core and the pool would thrash
    saved = {k: os.environ.get(k) for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS")}
    for k in saved:
        os.environ[k] = str(threads_per_worker)
    try:
        return _cpu_pool_run(ctx, workers, model_name, size, K, n_images, steps, warmup, threads_per_worker)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _cpu_pool_run(ctx, workers, model_name, size, K, n_images, steps, warmup, threads_per_worker):
    from concurrent.futures import ProcessPoolExecutor
    with ProcessPoolExecutor(workers, mp_context=ctx, initializer=_ref_worker_init,
                             initargs=(model_name, threads_per_worker)) as ex:
        for w in range(max(1, warmup)):
            list(ex.map(_ref_worker_task, [(10_000 + i, size, K) for i in range(workers)], timeout=600))
        t0 = time.perf_counter()
        parts = []
        for s in range(steps):
            parts += list(ex.map(_ref_worker_task, [(s * n_images + i, size, K) for i in range(n_images)], timeout=600))
        dt = time.perf_counter() - t0
    split = {"vit_s_per_image": sum(p[0] for p in parts) / len(parts), "eigs_s_per_image": sum(p[1] for p in parts) / len(parts)}
    return n_images * steps / dt, dt, {"workers": workers, "threads_per_worker": threads_per_worker, **split}


def run_reference(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n = args.ref_images_per_step if args.ref_images_per_step > 0 else 2 * max(1, min(cores // 2, 64))
    value, dt, split = cpu_pool_images_per_sec(args.model, args.size, args.K, n, args.steps, args.warmup)
    sample = (f"{n} synthetic {args.size}x{args.size} images per step (a bounded sample of the {args.images_per_step}-image step); "
              f"fp32 eager DINO ViT + the reference's scipy eigsh route in {split['workers']} worker processes x "
              f"{split['threads_per_worker']} BLAS threads")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"configs[1]: {args.size}x{args.size} {args.model} dense affinity K={args.K}",
                       "images_per_step": n},
            "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": "port",
                             "sample": sample, "split": split},
            "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------
def run_ours(args, rank, local_rank, world):
    import numpy as np
    import torch
    import torch.distributed as dist
    torch.set_grad_enabled(False)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: the hot path has no CPU fallback (use --impl reference for the CPU path)")
    _lib = importlib.import_module(PKG + "._lib")
    pipeline = importlib.import_module(PKG + ".pipeline")
    synth = importlib.import_module(PKG + ".synth")
    vit = importlib.import_module(PKG + ".vit")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        # NCCL prints its version banner (NCCL_DEBUG=VERSION in this image) to stdout by default; stdout carries the
        # ONE JSON line only
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    # the one collective of the path: DINO weights from rank 0
    sd0 = vit.random_state_dict(args.model, 0) if rank == 0 else None
    sd = pipeline.broadcast_weights(args.model, 0, device=dev, src=0, state_dict=sd0)
    if args.vit_batch <= 0:
        args.vit_batch = args.images_per_step
    pipe = pipeline.SpectralPipeline(args.model, K=args.K, device=dev, state_dict=sd, vit_batch=args.vit_batch)
    P, d, depth = pipe.model.patch_size, pipe.model.dim, pipe.model.depth
    B, S, K = args.images_per_step, args.size, args.K
    N = (S // P) ** 2
    T = N + 1
    host_imgs = synth.blobs_batch(B, S, S, seed0=rank * B).pin_memory()
    dev_imgs = host_imgs.to(dev)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        pipe.run_device(dev_imgs)
    barrier()
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark_begin()
    e0.record()
    info = None
    for _ in range(args.steps):
        _, _, info = pipe.run_device(dev_imgs)
    e1.record()
    barrier()
    sampler.mark_end()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = _lib.launch_count() - n0
    value = world * B * args.steps / (ms * 1e-3)
    conv = int(info[:, 1].sum().item())
    steps_mean = float(info[:, 0].float().mean().item())

    # ---- end to end through the host-buffer call: every step copies its uint8 images from pinned host memory and
    # brings the eigenvectors back; the streaming driver overlaps step i+1's H2D and step i-1's D2H with step i
    for _ in pipe.run_host_pipelined([host_imgs] * 2):
        pass
    barrier()
    sampler.mark_begin()
    e0.record()
    out = None
    for out in pipe.run_host_pipelined([host_imgs] * args.steps):
        pass
    e1.record()
    barrier()
    sampler.mark_end()
    clocks = sampler.stop()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    e2e_value = world * B * args.steps / (ms_e2e * 1e-3)
    h2d = int(host_imgs.numel())
    d2h = int(out[0].numel() * 4 + out[1].numel() * 4 + out[2].numel() * 4)

    # ---- instrumented pass: CUDA events around every launch of the library, same K steps
    _lib.profile(True)
    for _ in range(args.steps):
        pipe.run_device(dev_imgs)
    torch.cuda.synchronize()
    prof = _lib.profile_read()
    _lib.profile(False)
    peaks = load_peaks()
    Mtok = args.vit_batch * T
    hid = 4 * d
    n_vit = (B + args.vit_batch - 1) // args.vit_batch
    m_steps = steps_mean
    # algorithmic work per launch: GEMM/attention FLOPs, affinity/eigsh bytes
    alg = {
        "gemm_patch": ("tensor", 2.0 * args.vit_batch * N * (3 * P * P) * d),
        "gemm_qkv": ("tensor", 2.0 * Mtok * d * 3 * d),
        "gemm_proj": ("tensor", 2.0 * Mtok * d * d),
        "gemm_fc1": ("tensor", 2.0 * Mtok * d * hid),
        "gemm_fc2": ("tensor", 2.0 * Mtok * hid * d),
        "gemm_kproj": ("tensor", 2.0 * Mtok * d * d),
        "attention": ("tensor", 4.0 * args.vit_batch * (d // 64) * T * T * 64),
        "layernorm": ("hbm", Mtok * d * (4 + 2)),
        "im2col": ("hbm", args.vit_batch * (S * S * 3 + N * 3 * P * P * 2)),
        "rownorm": ("hbm", B * N * d * 8),
        "affinity": ("hbm", B * (4.0 * N * d + 4.0 * N * N)),
        "eigsh": ("hbm", B * (4.0 * N * N * (m_steps + 1) + 4.0 * K * N)),
    }
    kernels = []
    total_ms = sum(v[1] for v in prof.values()) or 1.0
    for name, (n_l, t_ms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
        ent = {"kernel": name, "launches": n_l, "total_ms": round(t_ms, 4), "share": round(t_ms / total_ms, 4)}
        if name in alg and n_l:
            bound, work = alg[name]
            avg_s = t_ms / n_l * 1e-3
            if bound == "tensor":
                ach = work / avg_s / 1e12
                peak = peaks["tflops_sustained"]
                ent.update({"bound": "tensor", "achieved": round(ach, 2), "peak": peak, "unit": "TFLOP/s",
                            "frac": round(ach / peak, 4)})
            else:
                ach = work / avg_s / 1e9
                peak = peaks["hbm_gbs"]
                ent.update({"bound": "hbm", "achieved": round(ach, 1), "peak": peak, "unit": "GB/s",
                            "frac": round(ach / peak, 4)})
        kernels.append(ent)
    dom = next((k for k in kernels if "bound" in k), None)
    roofline = None
    if dom:
        traffic = None
        try:   # dram__bytes_read.sum + dram__bytes_write.sum per launch of that kernel, from the committed ncu capture
            per_img = json.loads((ROOT / "profiles" / "r1_traffic.json").read_text())["per_image_bytes"].get(dom["kernel"])
            imgs_per_launch = B if dom["kernel"] in ("eigsh", "affinity", "rownorm") else args.vit_batch
            traffic = per_img * imgs_per_launch if per_img else None
        except Exception:
            pass
        roofline = {"kernel": dom["kernel"], "bound": dom["bound"], "achieved": dom["achieved"], "peak": dom["peak"],
                    "unit": dom["unit"], "frac": dom["frac"], "traffic": traffic, "share_of_step": dom["share"],
                    "peak_source": peaks["source"] + (", sustained bf16 GEMM" if dom["bound"] == "tensor" else ""),
                    "timing": "CUDA events around every launch of the class in a separate instrumented pass of the same steps"}

    line = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16 operands / f32 accumulate (ViT); f32 (affinity, eigensolver)",
            "data": "synthetic",
            "config": {"workload": f"configs[1]: {S}x{S} {args.model} dense affinity K={K}", "images_per_step_per_gpu": B,
                       "vit_batch": args.vit_batch, "patches": N,
                       "l2": f"inputs larger than L2: {h2d / 1e6:.0f} MB of uint8 images and {B * N * N * 4 / 1e6:.0f} MB of affinity matrices per step",
                       "weights": "random init (upstream recipe), NCCL broadcast from rank 0" if world > 1 else "random init (upstream recipe)"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "eigensolver": {"converged": conv, "of": B, "lanczos_steps_mean": steps_mean},
            "roofline": roofline, "kernels": kernels}

    # ---- CPU baseline + parity on a bounded sample (rank 0, N=1 only)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import eigs_ref
        cores = os.cpu_count() or 1
        n_cpu = args.cpu_sample if args.cpu_sample > 0 else 2 * max(1, min(cores // 2, 64))
        ips, _, split = cpu_pool_images_per_sec(args.model, S, K, n_cpu, 1, 1)
        line["cpu_baseline"] = {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                                "sample": f"{n_cpu} synthetic {S}x{S} images; fp32 eager DINO ViT + the reference's scipy eigsh route, "
                                          f"{split['workers']} worker processes x {split['threads_per_worker']} BLAS threads",
                                "split": split}
        n = min(args.parity_sample, B)
        torch.set_num_threads(cores)
        _, outs, _ = cpu_reference_images_per_sec(host_imgs[:n], args.model, K, sd, want_outputs=True)
        feats = pipe._bufs["feats"][:n].cpu()
        evecs = out[1][:n]
        worst_same = worst_abs = worst_e2e = 0.0
        for i in range(n):
            ev_o, vec_o = eigs_ref.extract_eig(feats[i], K)  # oracle on OUR features: the 1e-4 criterion
            v = evecs[i].numpy()
            for k in range(K):
                vo = vec_o[k].numpy()
                s = np.sign(np.dot(v[k], vo)) or 1.0
                worst_same = max(worst_same, float(np.linalg.norm(v[k] - s * vo) / np.linalg.norm(vo)))
                worst_abs = max(worst_abs, float(np.abs(v[k] - s * vo).max()))
                ve = outs[i][2][k].numpy()                  # oracle end to end (fp32 ViT on the CPU)
                s = np.sign(np.dot(v[k], ve)) or 1.0
                worst_e2e = max(worst_e2e, float(np.linalg.norm(v[k] - s * ve) / np.linalg.norm(ve)))
        line["parity"] = {"eigvec_max_abs_err": worst_abs, "eigvec_rel_l2_same_features": worst_same,
                          "eigvec_rel_l2_end_to_end_vs_fp32_oracle": worst_e2e, "images": n, "tolerance": 1e-4}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
