#!/usr/bin/env python
"""bench.py -- frames/sec of the StyleGAN3-T 1024^2 audio-reactive render (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU cores

One "step" = one batch of `--batch` frames per GPU through the synthesis hot path (latents resident in
HBM -> uint8 frames resident in HBM).  Frames are independent, so ranks render disjoint frame ranges
with no data-path collective (weights and latents are broadcast once before the timed region):
"scaling": "weak".  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "frames/sec StyleGAN3 1024^2 audio-reactive render"
WORKLOAD = "StyleGAN3-T 1024^2 random-init, 30 s @ 24 fps (720 frames) audio-reactive latents, 48 kHz sine sweep"


_REAL_STDOUT = None


def guard_stdout():
    """The driver reads ONE JSON line from stdout: route everything libraries print there (the NCCL version banner
    of a multi-rank run, warnings) to stderr and keep the real stdout for emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def measured_peaks():
    """HBM GB/s and dense bf16 TFLOP/s to divide by: the driver-written MEASURED_PEAKS.json (keys hbm_gbs, bf16_tflops,
    bf16_tflops_sustained: B200_PROFILING.md) when present and readable, else the fallback figures of that guide."""
    fallback = dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if not os.path.exists(p):
        return fallback
    try:
        d = json.load(open(p))
        burst = float(d["bf16_tflops"])
        return dict(hbm=float(d["hbm_gbs"]), tf_burst=burst, tf_sustained=float(d.get("bf16_tflops_sustained", burst)),
                    source="measured (MEASURED_PEAKS.json)")
    except (OSError, ValueError, KeyError, TypeError):
        return dict(fallback, source="fallback (B200_PROFILING.md; MEASURED_PEAKS.json unreadable)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                       "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.25)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, reasons, smax = [], set(), None
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); smax = float(r[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=smax, reasons=sorted(reasons), samples=len(sm))
        return out


def sg3_algorithmic_work(geo, B):
    """Per-step algorithmic work: modulated-conv FLOPs and filtered_lrelu HBM bytes (fp16)."""
    flops = 0.0
    fl_bytes = 0.0
    for g in geo["layers"]:
        hc = g["in_size"] + g["conv_kernel"] - 1
        flops += 2.0 * g["in_channels"] * g["out_channels"] * g["conv_kernel"] ** 2 * hc * hc
        if not g["is_torgb"]:
            fl_bytes += g["out_channels"] * (hc * hc + g["out_size"] ** 2) * 2.0
        else:
            fl_bytes += g["out_channels"] * (hc * hc + g["out_size"] ** 2) * 2.0
    return flops * B, fl_bytes * B


# ----------------------------------------------------------------------------------------------------
# reference arm: the reference algorithm (oracle restatement, see oracle/sg3.py header) on the host CPU
# ----------------------------------------------------------------------------------------------------
def run_reference(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import sg3 as O
    from maua_b200.workload import c2_latents

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    net = O.make_synthesis("T", 1024, seed=0)
    lat, _ = c2_latents(net.num_ws)
    # one step = ONE frame of the 720-frame job (a bounded sample: the CPU needs tens of seconds per frame);
    # the timed steps are additionally capped by a wall-clock budget so the arm ends within minutes.
    budget_s = float(os.environ.get("MB_REF_BUDGET_S", "200"))
    t_start = time.perf_counter()
    done_w = 0
    per = None
    for i in range(args.warmup):
        t0 = time.perf_counter()
        net(lat[i:i + 1])
        per = time.perf_counter() - t0
        done_w += 1
        if (time.perf_counter() - t_start) + 2 * per > budget_s * 0.4:
            break
    times = []
    for i in range(args.steps):
        t0 = time.perf_counter()
        net(lat[(done_w + i) % len(lat)][None])
        times.append(time.perf_counter() - t0)
        if (time.perf_counter() - t_start) + times[-1] > budget_s:
            break
    total = sum(times)
    fps = len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": len(times), "steps_requested": args.steps, "warmup": done_w, "ms_per_step": 1000 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step": 1},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": f"{len(times)} of 720 frames, one frame per step (fp32 PyTorch restatement of the "
                                   "reference algorithm; the reference's own network source is an un-vendored submodule)"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ----------------------------------------------------------------------------------------------------
# this repo's arm
# ----------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: maua_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from maua_b200.GAN.wrappers import get_generator_class
    from maua_b200.workload import c2_latents_device

    B, K, W = args.batch, args.steps, args.warmup
    torch.manual_seed(0)
    G = get_generator_class("stylegan3")(model_file=None).to(dev)
    net = G.synthesizer.G_synth
    if world > 1:  # generator weights broadcast once over NCCL
        for t in list(net.parameters()) + list(net.buffers()):
            dist.broadcast(t.data, src=0)
    # audio-reactive latents of the 720-frame job: built on rank 0, broadcast, sharded by contiguous frame range
    audio_info = {}
    if rank == 0:
        lat, audio_info = c2_latents_device(net.num_ws, dev)   # onset / rms features by the library's audio kernels
    else:
        lat = torch.empty(720, net.num_ws, 512, device=dev)
    if world > 1:
        dist.broadcast(lat, src=0)
    T = lat.shape[0]
    per = T // world
    my = lat[rank * per:(rank + 1) * per].contiguous()
    nb = max(per // B, 1)

    frames = torch.empty(B, 1024, 1024, 3, device=dev, dtype=torch.uint8)
    def step(i):
        j = (i % nb) * B
        net(my[j:j + B], out_fmt="u8", out=frames)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(W):
        step(i)
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    prof_acc = {}
    net.set_option("profile", 2)  # the library records CUDA events around every launch of the timed steps
    net.set_option("profile_reset", 1)
    barrier()
    e0.record()
    for i in range(K):
        step(W + i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    # per-kernel device times averaged over the K timed steps (events recorded on the launching stream)
    for kind, layer, t in net.profile_read():
        prof_acc.setdefault(kind, []).append((layer, t / K))
    clk = clocks.stop() if rank == 0 else None
    tmax = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax.item())
    launches = net.last_launch_count() * K

    # ---- e2e: the public call with HOST buffers; H2D of the latents and D2H of the frames every step ----
    from maua_b200.audiovisual.render._loop import AsyncFrameDownloader

    host_lat = my[: nb * B].cpu().pin_memory()
    dl = AsyncFrameDownloader((B, 1024, 1024, 3), dev, depth=2)   # pinned host ring, D2H on a side stream
    net.set_option("profile", 0)

    def e2e_step(i):
        j = (i % nb) * B
        ws = host_lat[j:j + B].to(dev, non_blocking=True)
        net(ws, out_fmt="u8", out=dl.device_buffer(i))
        dl.download(i)

    for i in range(min(W, 2)):
        e2e_step(i)
    dl.synchronize()
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        e2e_step(i)
    dl.synchronize()          # the last batch's frames are in pinned host memory
    torch.cuda.synchronize()
    t_e2e = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_fps = world * B * K / float(t_e2e.item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks = measured_peaks()
    geo = net.geometry
    flops, fl_bytes = sg3_algorithmic_work(geo, B)
    names = {0: "styles", 1: "input", 2: "modulated_conv2d (tcgen05)", 3: "filtered_lrelu", 4: "layout transpose", 5: "torgb+output"}
    per_kind = {k: sum(t for _, t in v) for k, v in prof_acc.items()}
    step_ms = sum(per_kind.values())
    dominant = max(per_kind, key=per_kind.get)
    fl_ms = per_kind.get(3, 0.0)
    conv_ms = per_kind.get(2, 0.0)
    roof_fl = {"kernel": "filtered_lrelu (14 launches / step)", "bound": "hbm", "achieved": fl_bytes / (fl_ms * 1e-3) / 1e9 if fl_ms else None,
               "peak": peaks["hbm"], "unit": "GB/s", "traffic": None, "peak_source": peaks["source"],
               "share_of_step": fl_ms / step_ms if step_ms else None, "ms_per_step": fl_ms}
    roof_fl["frac"] = roof_fl["achieved"] / roof_fl["peak"] if roof_fl["achieved"] else None
    roof_conv = {"kernel": "modulated_conv2d tcgen05 implicit GEMM (14 launches / step)", "bound": "tensor",
                 "achieved": flops / (conv_ms * 1e-3) / 1e12 if conv_ms else None, "peak": peaks["tf_sustained"],
                 "unit": "TFLOP/s", "traffic": None, "peak_source": peaks["source"] + ", sustained (kernel timed inside a long step)",
                 "share_of_step": conv_ms / step_ms if step_ms else None, "ms_per_step": conv_ms}
    roof_conv["frac"] = roof_conv["achieved"] / roof_conv["peak"] if roof_conv["achieved"] else None
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        tr = json.load(open(traffic_file))
        roof_fl["traffic"] = tr.get("filtered_lrelu_bytes_per_step_b%d" % B)
        roof_conv["traffic"] = tr.get("modulated_conv2d_bytes_per_step_b%d" % B)
        roof_fl["traffic_note"] = roof_conv["traffic_note"] = (
            "bytes per step: ncu dram__bytes_read.sum + dram__bytes_write.sum summed over the kernel family's 14 launches of one "
            "forward (profiles/traffic.json); algorithmic bytes per step = achieved * ms_per_step")

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import sg3 as O

        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        onet = O.make_synthesis("T", 1024, seed=0)
        wl = my[:1].cpu()
        t0 = time.perf_counter()
        onet(wl)
        dt = time.perf_counter() - t0
        cpu = {"value": 1.0 / dt, "unit": "frames/s", "cores": cores, "kind": "port",
               "sample": "1 of 720 frames (fp32 PyTorch restatement of the reference algorithm, all host threads)"}

    line = {
        "metric": METRIC, "value": world * B * K / (ms * 1e-3), "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step_per_gpu": B, "frames_per_gpu": per,
                   "l2": "per-step working set (GBs of activations) >> 126 MB L2, no explicit flush",
                   "output": "uint8 NHWC frames resident in HBM",
                   "audio_features": {**audio_info, "note": "device STFT/HPSS/onset/rms + chromagram (harmonic, tuning estimate, constant-Q, CENS) pass run once before the timed region"}, "sharding": f"contiguous frame ranges over {world} rank(s), no data-path collective"},
        "clocks": clk,
        "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": B * net.num_ws * 512 * 4,
                "d2h_bytes_per_step": B * 1024 * 1024 * 3},
        "gpu_launches": launches,
        "roofline": roof_fl if dominant == 3 else roof_conv,
        "roofline_other": roof_conv if dominant == 3 else roof_fl,
        "kernel_ms_per_step": {names[k]: round(v, 4) for k, v in sorted(per_kind.items())},
        "cpu_baseline": cpu,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=16,
                    help="frames per step per GPU (16 = the reference FFMPEG renderer's batch size, maua/audiovisual/render/ffmpeg.py:31)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        guard_stdout()
        return run_reference(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: `python bench.py --gpus N` re-launches itself under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29511"), os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    guard_stdout()
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
