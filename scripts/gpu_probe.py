"""First-contact GPU probe: runs each op once with verbose diagnostics (used through gpurun)."""
import sys, os, time, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import sg3 as O
from maua_b200 import ops

dev = torch.device("cuda:0")
IMPLS = (1, 0, 2)
print(torch.cuda.get_device_name(0), flush=True)


def rel(got, ref):
    ref = ref.double(); got = got.double().cpu()
    return float((got - ref).abs().max() / ref.square().mean().sqrt())


def conv_case(B, Cin, Cout, H, W, k, demod=True):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, Cin, H, W, generator=g); w = torch.randn(Cout, Cin, k, k, generator=g); s = torch.randn(B, Cin, generator=g) + 1
    ref = O.modulated_conv2d_ref(x, w, s, demodulate=demod, padding=k - 1)
    for impl in IMPLS:
        try:
            got = ops.modulated_conv2d(x.to(dev), w.to(dev), s.to(dev), demodulate=demod, impl=impl)
            torch.cuda.synchronize()
            e = rel(got, ref)
            print(f"conv B{B} Cin{Cin} Cout{Cout} {H}x{W} k{k} impl{impl}: rel err {e:.3e}", flush=True)
            if e > 1e-2:
                d = (got.cpu() - ref).abs()
                print("   worst per-channel err:", d.amax(dim=(0, 2, 3))[:8].tolist())
                print("   err rows:", d.amax(dim=(0, 1, 3))[:12].tolist())
                print("   err cols:", d.amax(dim=(0, 1, 2))[:12].tolist())
                print("   got[0,0,:3,:6]", got[0, 0, :3, :6].cpu().tolist()); print("   ref[0,0,:3,:6]", ref[0, 0, :3, :6].tolist())
        except Exception as ex:
            print(f"conv impl{impl} FAILED: {str(ex)[:200]}", flush=True)
            from maua_b200 import _lib
            print("   debug words:", _lib.debug_words(), flush=True)


def fl_case(C, H, W, up, down, ut, dt, lo, hi, radial=False):
    g = torch.Generator().manual_seed(2)
    x = (torch.randn(2, C, H, W, generator=g) * 2).half().float(); b = torch.randn(C, generator=g)
    fu = O.design_lowpass_filter(ut, 8.0, 9.0, 64.0) if ut > 1 else None
    fd = O.design_lowpass_filter(dt, 8.0, 9.0, 64.0, radial=radial) if dt > 1 else None
    pad = [lo, hi, lo, hi]
    ref = O.filtered_lrelu_ref(x, fu=fu, fd=fd, b=b, up=up, down=down, padding=pad, gain=np.sqrt(2), slope=0.2, clamp=256)
    for impl in ("1", "2", "0"):
        os.environ["MB_FLRELU_IMPL"] = impl
        try:
            got = ops.filtered_lrelu(x.to(dev), None if fu is None else fu.to(dev), None if fd is None else fd.to(dev), b.to(dev),
                                     up=up, down=down, padding=pad, gain=np.sqrt(2), slope=0.2, clamp=256)
            print(f"flrelu C{C} {H}x{W} up{up} down{down} pad({lo},{hi}) radial{radial} impl{impl}: shape {tuple(got.shape)} vs {tuple(ref.shape)} rel err {rel(got, ref):.3e}", flush=True)
        except Exception as ex:
            print(f"flrelu impl{impl} FAILED: {ex}", flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which.startswith("conv") and len(which) > 4:
        IMPLS = (int(which[4:]),)
        which = "conv"
    if which in ("all", "fl"):
        fl_case(3, 38, 38, 2, 2, 12, 12, 9, 8)
        fl_case(2, 54, 54, 4, 2, 24, 12, -6, -9)
        fl_case(2, 150, 130, 2, 2, 12, 12, -11, -12)
        fl_case(2, 86, 86, 4, 2, 24, 12, -6, -9, True)
        fl_case(4, 33, 47, 1, 1, 1, 1, 0, 0)
        fl_case(1, 200, 180, 4, 2, 24, 12, -7, -8)
        fl_case(1, 90, 70, 2, 2, 12, 12, -10, -13)
    if which in ("all", "conv"):
        conv_case(1, 64, 128, 20, 20, 3)
        conv_case(2, 81, 51, 70, 66, 3)
        conv_case(2, 128, 96, 40, 24, 1)
        conv_case(1, 512, 512, 38, 38, 3)
        conv_case(2, 51, 32, 100, 90, 3)
        conv_case(1, 32, 32, 70, 70, 3)
        conv_case(1, 32, 3, 64, 64, 1, False)
    print("probe done", flush=True)


def net_case(res=256, cb=8192, cm=128, B=2, config="T"):
    from maua_b200.GAN.networks import stylegan3 as N
    kw = dict(channel_base=cb, channel_max=cm)
    if config == "R":
        kw.update(conv_kernel=1, use_radial_filters=True)
    onet = O.make_synthesis("T", img_resolution=res, seed=0, **kw)
    torch.manual_seed(0)
    net = N.SynthesisNetwork(w_dim=512, img_resolution=res, img_channels=3, **kw)
    net.load_state_dict(onet.state_dict())
    torch.manual_seed(5)
    ws = torch.randn(B, net.num_ws, 512)
    ref, acts = onet(ws, return_activations=True)
    layers = [getattr(onet, n) for n in onet.layer_names]
    wsu = ws.unbind(1)

    def style(i):
        l = layers[i]
        s = l.affine(wsu[i + 1])
        if l.is_torgb:
            return s * (1 / np.sqrt(l.in_channels * l.conv_kernel ** 2))
        return s * s.square().mean(1, keepdim=True).rsqrt()

    for impl in (1, 0):
        net.set_option("conv_impl", impl)
        for stop in range(-1, len(layers) - 1):
            net.set_option("debug_stop", stop)
            try:
                net(ws.to(dev))
                a = net.read_activation(B)
                torch.cuda.synchronize()
                want = acts[stop + 1] * style(stop + 1)[:, :, None, None]
                print(f"net{res} impl{impl} after layer {stop}: shape {tuple(a.shape)} rel err {rel(a, want):.3e}", flush=True)
            except Exception as ex:
                print(f"net impl{impl} stop{stop} FAILED: {ex}", flush=True)
                break
        net.set_option("debug_stop", 1 << 30)
        try:
            out = net(ws.to(dev))
            torch.cuda.synchronize()
            pe = ((out.cpu() + 1) / 2).clamp(0, 1) - ((ref + 1) / 2).clamp(0, 1)
            print(f"net{res} impl{impl} FINAL: rel {rel(out, ref):.3e} max-abs pixel err {pe.abs().max():.3e}  launches {net.last_launch_count()}", flush=True)
            u8 = net(ws.to(dev), out_fmt="u8")
            want8 = (((ref + 1) / 2).clamp(0, 1) * 255).round().permute(0, 2, 3, 1)
            print("   u8 max diff", (u8.cpu().float() - want8).abs().max().item(), flush=True)
        except Exception as ex:
            print(f"net impl{impl} final FAILED: {ex}", flush=True)


if __name__ == "__main__" and (len(sys.argv) > 1 and sys.argv[1] in ("net", "allnet")):
    net_case()
    print("net probe done", flush=True)
