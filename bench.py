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
# --config: BASELINE.json configs[1] (the default, the configuration the metric is quoted on) and configs[2]
CONFIGS = {
    "c2": dict(arch="T", seconds=30.0, fps=24, workload=WORKLOAD),
    "c3": dict(arch="R", seconds=180.0, fps=60,
               workload="StyleGAN3-R 1024^2 random-init, 180 s @ 60 fps (10 800 frames) audio-reactive latents, 48 kHz sine sweep, "
                        "frames sharded over the ranks"),
}


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
    cfg = CONFIGS.get(args.config, CONFIGS["c2"])
    net = O.make_synthesis(cfg["arch"], 1024, seed=0)
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
        "config": {"workload": cfg["workload"], "name": args.config, "frames_per_step": 1},
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
    from maua_b200.workload import job_latents_device, resampled_sweep, audio_reactive_latents

    B, K, W = args.batch, args.steps, args.warmup
    cfg = CONFIGS[args.config]
    n_job = int(round(cfg["seconds"] * cfg["fps"]))
    torch.manual_seed(0)
    G = get_generator_class("stylegan3")(model_file=None)
    if cfg["arch"] == "R":
        from maua_b200.GAN.networks import stylegan3 as N

        torch.manual_seed(0)
        G.synthesizer.G_synth = N.SynthesisNetwork(w_dim=512, img_resolution=1024, img_channels=3, **N.SG3_R_KWARGS)
    G = G.to(dev)
    net = G.synthesizer.G_synth
    if world > 1:  # generator weights broadcast once over NCCL
        for t in list(net.parameters()) + list(net.buffers()):
            dist.broadcast(t.data, src=0)
    # audio-reactive latents of the job: built on rank 0, broadcast, sharded by contiguous frame range
    audio_info = {}
    if rank == 0:
        lat, audio_info = job_latents_device(net.num_ws, dev, cfg["seconds"], cfg["fps"])   # the library's audio kernels
    else:
        lat = torch.empty(n_job, net.num_ws, 512, device=dev)
    if world > 1:
        dist.broadcast(lat, src=0)
    T = lat.shape[0]
    per = T // world
    my = lat[rank * per:(rank + 1) * per].contiguous()
    nb = max(per // B, 1)

    # Finished uint8 frames are gathered to rank 0 over NCCL INSIDE the timed region (north_star: "rendered frames gathered
    # across the GPUs with NCCL over NVLink"): one gather per step (grouped ncclSend / ncclRecv on NCCL's own stream),
    # double-buffered so the transfer of step i overlaps the kernels of step i + 1.
    frame_shape = (B, 1024, 1024, 3)
    frames = [torch.empty(frame_shape, device=dev, dtype=torch.uint8) for _ in range(2)]
    gathered = [[torch.empty(frame_shape, device=dev, dtype=torch.uint8) for _ in range(world)] for _ in range(2)] if (world > 1 and rank == 0) else None
    pending = [None, None]
    gather_bytes = (world - 1) * B * 1024 * 1024 * 3 if world > 1 else 0

    def step(i):
        j = (i % nb) * B
        k = i % 2
        if pending[k] is not None:
            pending[k].wait()          # stream-level: the gather that read frames[k] two steps ago is done
        net(my[j:j + B], out_fmt="u8", out=frames[k])
        if world > 1:
            pending[k] = dist.gather(frames[k], gathered[k] if rank == 0 else None, dst=0, async_op=True)

    def drain():
        for k in range(2):
            if pending[k] is not None:
                pending[k].wait()
                pending[k] = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(W):
        step(i)
    drain()
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
    drain()                      # the last gathers are inside the timed region
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

    # ---- e2e: the product's own entry points with HOST buffers ------------------------------------------------------------
    # What a reference-style patch + renderer does for this rank's frames, timed wall-clock from host audio to the last byte in
    # the sink: (1) the track (pinned host memory) -> device -> resample -> audio features -> latent sequence -> host tensors
    # (process_audio / process_synthesizer_inputs of a patch), (2) FFMPEG.__call__(synthesizer, {"latents": host}, postprocess)
    # = pinned staging, per-batch H2D, synthesis, (x+1)/2, postprocess, tensor2bytes, D2H into the pinned ring, writer thread
    # -> sink (a byte counter standing in for the ffmpeg pipe: the x264 encode is not part of the path).
    from maua_b200.audiovisual.render.ffmpeg import FFMPEG
    from maua_b200.workload import sine_sweep

    net.set_option("profile", 0)

    class ByteCounter:
        def __init__(self):
            self.n = 0

        def write(self, view):
            self.n += len(view)

        def close(self):
            pass

    track48, sr48 = sine_sweep(cfg["seconds"], tremolo_hz=4.0)
    track_host = torch.from_numpy(track48).pin_memory()
    # every rank renders a whole 720-frame job (configs[1]'s length; its shard's latents, cycled when the shard is shorter), so that
    # the one-off costs of a render (audio pass, pipeline fill and drain) weigh what they weigh in the real job, whatever --steps is
    n_e2e = 720
    e2e_index = (torch.arange(n_e2e, device=dev) % per) + rank * per
    postprocess = lambda video: video   # noqa: E731  (a patch's process_outputs + force_output_size at native size)
    postprocess.pure = True              # what generate_audiovisal_from_patch declares for a patch's stock stages

    host_stage = {}

    def e2e_job(n_frames):
        import torchaudio

        y48 = track_host.to(dev, non_blocking=True)
        sr = 1024 * cfg["fps"]
        y = torchaudio.functional.resample(y48, sr48, sr)
        n = n_job * 1024
        y = (y[:n] if y.numel() >= n else torch.nn.functional.pad(y, (0, n - y.numel()))).contiguous()
        lat_dev = audio_reactive_latents(y, sr, net.num_ws)
        t_audio = time.perf_counter()
        if "buf" not in host_stage:   # the application's reusable pinned staging buffer for the latent sequence (allocated once)
            host_stage["buf"] = torch.empty((n_e2e,) + tuple(lat_dev.shape[1:]), dtype=lat_dev.dtype).pin_memory()
        host_lat = host_stage["buf"][:n_frames]
        host_lat.copy_(lat_dev[e2e_index[:n_frames]])                               # the patch hands host tensors to the renderer
        sink = ByteCounter()
        FFMPEG(None, fps=cfg["fps"], batch_size=B, sink=sink)(G.synthesizer, {"latents": host_lat}, postprocess)
        return sink.n, t_audio

    e2e_job(3 * B)               # warm-up of this path (pinned rings, resample kernel cache)
    barrier()
    t0 = time.perf_counter()
    nbytes, t_audio = e2e_job(n_e2e)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    assert nbytes == n_e2e * 1024 * 1024 * 3, (nbytes, n_e2e)
    t_e2e = torch.tensor([t1 - t0], device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_fps = world * n_e2e / float(t_e2e.item())
    e2e_audio_ms = 1000.0 * (t_audio - t0)

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
        tag = "" if cfg["arch"] == "T" else cfg["arch"] + "_"
        roof_fl["traffic"] = tr.get("filtered_lrelu_bytes_per_step_%sb%d" % (tag, B))
        roof_conv["traffic"] = tr.get("modulated_conv2d_bytes_per_step_%sb%d" % (tag, B))
        roof_fl["traffic_note"] = roof_conv["traffic_note"] = (
            "bytes per step: ncu dram__bytes_read.sum + dram__bytes_write.sum summed over the kernel family's 14 launches of one "
            "forward (profiles/traffic.json); algorithmic bytes per step = achieved * ms_per_step")

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import sg3 as O

        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        onet = O.make_synthesis(cfg["arch"], 1024, seed=0)
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
        "config": {"workload": cfg["workload"], "name": args.config, "frames_per_step_per_gpu": B, "frames_per_gpu": per,
                   "collective": (f"per step: gather of the step's uint8 frames to rank 0 over NCCL (grouped ncclSend/ncclRecv, "
                                  f"{gather_bytes} bytes into the root per step), overlapped with the next step's kernels; weights and "
                                  f"latents broadcast once before the timed region") if world > 1 else "none (one rank)",
                   "l2": "per-step working set (GBs of activations) >> 126 MB L2, no explicit flush",
                   "output": "uint8 NHWC frames resident in HBM",
                   "audio_features": {**audio_info, "note": "device STFT/HPSS/onset/rms + chromagram (harmonic, tuning estimate, constant-Q, CENS) pass run once before the timed region"}, "sharding": f"contiguous frame ranges over {world} rank(s)"},
        "clocks": clk,
        "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": B * net.num_ws * 512 * 4,
                "d2h_bytes_per_step": B * 1024 * 1024 * 3,
                "path": "host audio -> device features + latent sequence -> host latents -> FFMPEG.__call__ (pinned staging, H2D per "
                        "batch, synthesis, postprocess, tensor2bytes, D2H ring, writer thread -> byte-counting sink)",
                "frames_per_gpu": n_e2e, "audio_and_latents_ms": round(e2e_audio_ms, 2),
                "h2d_bytes_once": int(track_host.numel() * 4)},
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


# ----------------------------------------------------------------------------------------------------
# secondary line: the StyleGAN2 path (the generator of selfsupervised/sample.generate), --config sg2
# ----------------------------------------------------------------------------------------------------
def sg2_algorithmic_work(res, channel_base=32768, channel_max=512, img_channels=3):
    """Per-frame algorithmic work of the in-tree StyleGAN2 synthesis network: modulated-conv FLOPs
    (2 Cin Cout 9 h_out^2 for the 3x3 convs, 2 Cin Cout 9 h_in^2 for the stride-2 transposed conv, + ToRGB) and the bytes the
    fused FIR / bias_act kernel has to move once (conv output in, next conv input out, fp16 pairs in `precise` mode)."""
    ch = lambda r: min(channel_base // r, channel_max)   # noqa: E731
    flops, act_vals = 0.0, 0.0
    r = 4
    while r <= res:
        co = ch(r)
        if r > 4:
            ci = ch(r // 2)
            flops += 2.0 * ci * co * 9 * (r // 2) ** 2            # transposed conv: one MAC set per INPUT pixel
            act_vals += co * ((r + 1) ** 2 + r * r)                # FIR input (2h+1)^2, activation output
        flops += 2.0 * co * co * 9 * r * r
        act_vals += co * 2 * r * r
        flops += 2.0 * co * img_channels * r * r
        r *= 2
    return flops, act_vals


def run_sg2(args):
    import torch

    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        raise RuntimeError("--config sg2 is a single-GPU line")
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: maua_b200 has no CPU fallback")
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    from maua_b200.GAN.networks import stylegan2 as N2
    from maua_b200.GAN.wrappers.stylegan2 import StyleGAN2Synthesizer
    from maua_b200.audiovisual.render.ffmpeg import FFMPEG
    from maua_b200.workload import job_latents_device

    B, K, W = min(args.batch, 8), args.steps, args.warmup
    res = 1024
    torch.manual_seed(0)
    net = N2.SynthesisNetwork(w_dim=512, img_resolution=res, img_channels=3).to(dev)
    S = StyleGAN2Synthesizer.__new__(StyleGAN2Synthesizer)
    torch.nn.Module.__init__(S)
    S.G_synth, S._hook_handles, S._warp_hooks = net, [], {}
    lat, audio_info = job_latents_device(net.num_ws, dev, 30.0, 24)
    nb = lat.shape[0] // B
    frames = torch.empty(B, res, res, 3, device=dev, dtype=torch.uint8)

    def step(i):
        j = (i % nb) * B
        net(lat[j:j + B], out_fmt="u8", out=frames)

    for i in range(W):
        step(i)
    torch.cuda.synchronize()
    clocks = ClockSampler(0)
    clocks.start()
    net.set_option("profile", 2)
    net.set_option("profile_reset", 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(K):
        step(W + i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    per_kind = {}
    for kind, _, t in net.profile_read():
        per_kind[kind] = per_kind.get(kind, 0.0) + t / K
    clk = clocks.stop()
    net.set_option("profile", 0)
    launches = net.last_launch_count() * K

    class ByteCounter:
        n = 0

        def write(self, view):
            self.n += len(view)

        def close(self):
            pass

    def e2e_job(n_frames):
        host = torch.empty((n_frames,) + tuple(lat.shape[1:])).pin_memory()
        host.copy_(lat[:n_frames])
        sink = ByteCounter()
        FFMPEG(None, fps=24, batch_size=B, sink=sink)(S, {"latents": host}, lambda v: v)
        return sink.n

    e2e_job(3 * B)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n_e2e = lat.shape[0]
    nbytes = e2e_job(n_e2e)
    torch.cuda.synchronize()
    e2e_fps = n_e2e / (time.perf_counter() - t0)
    assert nbytes == n_e2e * res * res * 3

    peaks = measured_peaks()
    flops, act_vals = sg2_algorithmic_work(res)
    conv_ms, act_ms = per_kind.get(2, 0.0), per_kind.get(3, 0.0)
    bytes_per_val_in, bytes_per_val_out = 4.0, 6.0     # conv output hi + lo planes in; next conv input [xh | xl | xh] out
    act_bytes = act_vals / 2 * bytes_per_val_in + act_vals / 2 * bytes_per_val_out
    roof_conv = {"kernel": "modulated_conv2d tcgen05 implicit GEMM (fp16 hi+lo operands: 3 MMAs per algorithmic MAC)", "bound": "tensor",
                 "achieved": flops * B / (conv_ms * 1e-3) / 1e12 if conv_ms else None, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                 "traffic": None, "peak_source": peaks["source"] + ", sustained", "ms_per_step": conv_ms}
    roof_conv["frac"] = roof_conv["achieved"] / roof_conv["peak"] if roof_conv["achieved"] else None
    roof_act = {"kernel": "fused upfirdn2d + noise + bias_act + ToRGB kernel", "bound": "hbm",
                "achieved": act_bytes * B / (act_ms * 1e-3) / 1e9 if act_ms else None, "peak": peaks["hbm"], "unit": "GB/s", "traffic": None,
                "peak_source": peaks["source"], "ms_per_step": act_ms}
    roof_act["frac"] = roof_act["achieved"] / roof_act["peak"] if roof_act["achieved"] else None
    cpu = None
    if not args.no_cpu_baseline:
        from oracle import sg2 as O2

        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        onet = O2.make_synthesis(res, seed=0)
        t0 = time.perf_counter()
        onet(lat[:1].cpu())
        cpu = {"value": 1.0 / (time.perf_counter() - t0), "unit": "frames/s", "cores": cores, "kind": "port",
               "sample": "1 of 720 frames (oracle/sg2.py, pinned against the reference's in-tree inference network, all host threads)"}
    names = {0: "styles", 1: "const input", 2: "modulated_conv2d (tcgen05)", 3: "upfirdn2d+bias_act+ToRGB (fused)", 4: "feature warp", 5: "skip image / output"}
    emit({
        "metric": "frames/sec StyleGAN2 1024^2 audio-reactive render", "value": B * K / (ms * 1e-3), "unit": "frames/s", "n_gpus": 1,
        "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16 hi+lo pairs (fp32-class), fp32 accumulate", "data": "synthetic",
        "config": {"workload": "StyleGAN2 1024^2 random-init (the generator of selfsupervised/sample.generate), 30 s @ 24 fps audio-reactive "
                               "latents, constant noise", "name": "sg2", "frames_per_step_per_gpu": B, "audio_features": audio_info,
                   "l2": "per-step working set >> 126 MB L2, no explicit flush"},
        "clocks": clk,
        "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": B * net.num_ws * 512 * 4, "d2h_bytes_per_step": B * res * res * 3,
                "path": "host latents -> FFMPEG.__call__ -> byte-counting sink", "frames_per_gpu": n_e2e},
        "gpu_launches": launches,
        "roofline": roof_conv if conv_ms >= act_ms else roof_act, "roofline_other": roof_act if conv_ms >= act_ms else roof_conv,
        "kernel_ms_per_step": {names[k]: round(v, 4) for k, v in sorted(per_kind.items())},
        "cpu_baseline": cpu,
    })
    return 0


# ----------------------------------------------------------------------------------------------------
# BASELINE configs[4]: StyleGAN2 512^2 + RealESRGAN 4x fused upscale pipeline, 60 s @ 30 fps, --config c5
# ----------------------------------------------------------------------------------------------------
def rrdb_algorithmic_flops(h, w, num_block=23):
    """FLOPs per frame of RRDBNet(num_feat=64, num_grow_ch=32, scale=4) on an h x w input: 2 Cin Cout 9 per output pixel."""
    rdb = sum(2.0 * (64 + 32 * k) * (32 if k < 4 else 64) * 9 for k in range(5))
    per_px = 2.0 * 3 * 64 * 9 + num_block * 3 * rdb + 2.0 * 64 * 64 * 9
    return h * w * per_px + 4 * h * w * 2.0 * 64 * 64 * 9 + 16 * h * w * (2 * 2.0 * 64 * 64 * 9 + 2.0 * 64 * 3 * 9)


def run_c5(args):
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: maua_b200 has no CPU fallback")
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from maua_b200.GAN.networks import stylegan2 as N2
    from maua_b200.audiovisual.render._loop import AsyncFrameDownloader
    from maua_b200.super.image.models.realesrgan import RRDBNet
    from maua_b200.workload import job_latents_device

    B, K, W = min(args.batch, 4), args.steps, args.warmup
    res, fps, seconds = 512, 30, 60.0
    torch.manual_seed(0)
    g = N2.SynthesisNetwork(w_dim=512, img_resolution=res, img_channels=3).to(dev)
    up = RRDBNet(num_block=23).to(dev)
    if world > 1:   # weights broadcast once
        for t in list(g.parameters()) + list(g.buffers()) + list(up.parameters()):
            dist.broadcast(t.data, src=0)
    lat, audio_info = job_latents_device(g.num_ws, dev, seconds, fps)
    if world > 1:
        dist.broadcast(lat, src=0)
    per = lat.shape[0] // world
    my = lat[rank * per:(rank + 1) * per].contiguous()
    nb = max(per // B, 1)
    small = torch.empty(B, res, res, 3, device=dev, dtype=torch.uint8)
    big = [torch.empty(B, 4 * res, 4 * res, 3, device=dev, dtype=torch.uint8) for _ in range(2)]
    gathered = [[torch.empty_like(big[0]) for _ in range(world)] for _ in range(2)] if (world > 1 and rank == 0) else None
    pending = [None, None]
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]

    def step(i, marks=None):
        j, k = (i % nb) * B, i % 2
        if pending[k] is not None:
            pending[k].wait()
        if marks: marks[0].record()
        g(my[j:j + B], out_fmt="u8", out=small)                 # StyleGAN2 512^2 -> rgb24 frames in HBM
        if marks: marks[1].record()
        up(small, out_fmt="u8", out=big[k])                       # RealESRGAN x4 on the frames where they lie
        if marks: marks[2].record()
        if world > 1:
            pending[k] = dist.gather(big[k], gathered[k] if rank == 0 else None, dst=0, async_op=True)

    def drain():
        for k in range(2):
            if pending[k] is not None:
                pending[k].wait()
                pending[k] = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(W):
        step(i)
    drain()
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(K):
        step(W + i, ev[i])
    drain()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    sg2_ms = sum(m[0].elapsed_time(m[1]) for m in ev) / K
    up_ms = sum(m[1].elapsed_time(m[2]) for m in ev) / K
    clk = clocks.stop() if rank == 0 else None
    tmax = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax.item())
    launches = (g.last_launch_count() + up.last_launch_count()) * K

    # e2e: pinned host latents -> H2D per batch -> StyleGAN2 -> RealESRGAN -> rgb24 frames downloaded into a pinned ring
    n_e2e = min(per - per % B, 40 * B)
    host = torch.empty((n_e2e,) + tuple(my.shape[1:])).pin_memory()
    host.copy_(my[:n_e2e])
    dl = AsyncFrameDownloader((B, 4 * res, 4 * res, 3), dev, depth=2)

    def e2e_job(n):
        got = 0
        for bi, j in enumerate(range(0, n, B)):
            ws = host[j:j + B].to(dev, non_blocking=True)
            g(ws, out_fmt="u8", out=small)
            up(small, out_fmt="u8", out=dl.device_buffer(bi))
            dl.download(bi)
            if bi >= 1:
                got += dl.host(bi - 1).numel()
        got += dl.host((n - 1) // B).numel()
        return got

    e2e_job(2 * B)
    barrier()
    t0 = time.perf_counter()
    nbytes = e2e_job(n_e2e)
    torch.cuda.synchronize()
    t_e2e = torch.tensor([time.perf_counter() - t0], device=dev)
    assert nbytes == n_e2e * 16 * res * res * 3
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_fps = world * n_e2e / float(t_e2e.item())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks = measured_peaks()
    flops = rrdb_algorithmic_flops(res, res)
    roof = {"kernel": "RRDBNet 3x3 convs: stacked pixel-major tcgen05 tile, channels-last lrelu / residual epilogue (353 launches / step)",
            "bound": "tensor", "achieved": flops * B / (up_ms * 1e-3) / 1e12, "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "traffic": None,
            "peak_source": peaks["source"] + ", sustained", "ms_per_step": up_ms, "share_of_step": up_ms / (up_ms + sg2_ms)}
    roof["frac"] = roof["achieved"] / roof["peak"]
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import rrdb as OR, sg2 as O2

        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        onet = O2.make_synthesis(res, seed=0)
        t0 = time.perf_counter()
        img = onet(my[:1].cpu())
        t_sg2 = time.perf_counter() - t0
        ornet = OR.make(num_block=23)
        crop = ((img[:, :, :128, :128] + 1) / 2).clamp(0, 1)
        t0 = time.perf_counter()
        ornet(crop)
        t_up = (time.perf_counter() - t0) * 16.0
        cpu = {"value": 1.0 / (t_sg2 + t_up), "unit": "frames/s", "cores": cores, "kind": "port",
               "sample": "1 frame: StyleGAN2 512^2 oracle in full, RRDBNet oracle on a 128 x 128 crop (1/16 of the frame) scaled by 16; "
                         "fp32 PyTorch restatements, all host threads"}
    emit({
        "metric": "frames/sec StyleGAN2 512^2 + RealESRGAN x4 (2048^2 out) audio-reactive render", "value": world * B * K / (ms * 1e-3),
        "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f16 (StyleGAN2: hi+lo pairs), fp32 accumulate", "data": "synthetic",
        "config": {"workload": "StyleGAN2 512^2 random-init + RealESRGAN x4plus (RRDBNet, 23 blocks, random init) fused on the device, 60 s @ 30 fps "
                               "(1 800 frames) audio-reactive latents, frames sharded over the ranks", "name": "c5", "frames_per_step_per_gpu": B,
                   "frames_per_gpu": per, "audio_features": audio_info,
                   "collective": (f"per step: gather of the step's 2048^2 rgb24 frames to rank 0 over NCCL ({(world - 1) * B * 16 * res * res * 3} bytes "
                                  f"into the root per step), overlapped with the next step") if world > 1 else "none (one rank)",
                   "l2": "per-step working set (GBs of activations) >> 126 MB L2, no explicit flush"},
        "clocks": clk,
        "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": B * g.num_ws * 512 * 4, "d2h_bytes_per_step": B * 16 * res * res * 3,
                "path": "pinned host latents -> H2D per batch -> StyleGAN2 (u8) -> RRDBNet (u8 in / out) -> D2H into a pinned ring", "frames_per_gpu": n_e2e},
        "gpu_launches": launches,
        "roofline": roof,
        "kernel_ms_per_step": {"stylegan2 512^2 (all kernels)": round(sg2_ms, 4), "RRDBNet x4 (all kernels)": round(up_ms, 4)},
        "cpu_baseline": cpu,
    })
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
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS) + ["sg2", "c5"],
                    help="c2 = BASELINE configs[1] (default, the headline), c3 = configs[2], sg2 = the StyleGAN2-1024 path (secondary line), "
                         "c5 = configs[4] (StyleGAN2 512^2 + RealESRGAN x4)")
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
    if args.config == "sg2":
        return run_sg2(args)
    if args.config == "c5":
        return run_c5(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
