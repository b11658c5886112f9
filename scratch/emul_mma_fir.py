"""Emulate the tensor-core FIR chain: fp16 filter taps and fp16 rounding between the four separable passes."""
import torch, numpy as np, sys, time, os
import torch.nn.functional as F
from oracle import sg3
torch.set_num_threads(os.cpu_count())
def h(x): return x.half().float()

def qfilt(f, phases):
    # fp16 taps, then rescale each polyphase branch so its DC gain matches the exact filter (applied in fp32 outside the MMA)
    fq = h(f)
    corr = torch.ones_like(f)
    n = f.numel()
    for p in range(phases):
        idx = torch.arange(p, n, phases)
        corr[idx] = f[idx].sum() / fq[idx].sum()
    return fq, corr

DC = False
def fir_chain(y, l, filt_half, inter_half, split=False):
    # y: conv output fp32 [B,C,H,W]; returns filtered_lrelu result following reference semantics
    B, C, H, W = y.shape
    up, down = l.up_factor, l.down_factor
    fu, fd = l.up_filter, l.down_filter
    px0, px1, py0, py1 = l.padding
    x = y + l.bias.reshape(1, -1, 1, 1)
    if inter_half: x = h(x)   # smem staging of the input tile in fp16
    gain = 1.0 if l.is_torgb else float(np.sqrt(2)); slope = 1.0 if l.is_torgb else 0.2
    if fu is None:
        t = x
    else:
        fuq = fu * up
        if filt_half and not split:
            fq, corr = qfilt(fuq, up)
            fuq = fq * corr if DC else fq
        # zero insert
        xz = x.reshape(B, C, H, 1, W, 1); xz = F.pad(xz, [0, up - 1, 0, 0, 0, up - 1]).reshape(B, C, H * up, W * up)
        xz = F.pad(xz, [max(px0, 0), max(px1, 0), max(py0, 0), max(py1, 0)])
        xz = xz[:, :, max(-py0, 0): xz.shape[2] - max(-py1, 0), max(-px0, 0): xz.shape[3] - max(-px1, 0)]
        f = fuq.flip(0)[None, None].repeat(C, 1, 1)
        t = F.conv2d(xz, f.unsqueeze(2), groups=C)       # horizontal
        if inter_half: t = h(t)
        t = F.conv2d(t, f.unsqueeze(3), groups=C)        # vertical
    t = F.leaky_relu(t, slope) * gain
    t = t.clamp(-l.conv_clamp, l.conv_clamp)
    if inter_half: t = h(t)
    if fd is None:
        return t
    if filt_half and not split:
        fq, corr = qfilt(fd, 1)
        fdq = fq * corr if DC else fq
    else:
        fdq = fd
    f = fdq.flip(0)[None, None].repeat(C, 1, 1)
    o = F.conv2d(t, f.unsqueeze(2), groups=C)
    if inter_half: o = h(o)
    o = F.conv2d(o, f.unsqueeze(3), groups=C)
    return o[:, :, ::down, ::down]

def emulate(net, ws, filt_half=True, inter_half=True, split=False):
    ws_ = ws.float().unbind(1)
    x = net.input(ws_[0])
    layers = [getattr(net, n) for n in net.layer_names]
    def styles(l, w):
        s = l.affine(w)
        if l.is_torgb: s = s * (1 / np.sqrt(l.in_channels * l.conv_kernel ** 2))
        else: s = s * s.square().mean(1, keepdim=True).rsqrt()
        return s
    Sn = [styles(l, w) for l, w in zip(layers, ws_[1:])]
    xs = h(x * Sn[0][:, :, None, None])
    for i, l in enumerate(layers):
        W = l.weight
        if not l.is_torgb:
            Wn = W * W.square().mean([1, 2, 3], keepdim=True).rsqrt()
            d = (Sn[i].square() @ Wn.square().sum([2, 3]).t() + 1e-8).rsqrt()
            y = h(F.conv2d(xs, h(Wn), padding=l.conv_kernel - 1) * d[:, :, None, None])
            x = fir_chain(y, l, filt_half, inter_half, split)
        else:
            y = F.conv2d(xs, W, padding=0)
            x = (y + l.bias.reshape(1, -1, 1, 1)).clamp(-l.conv_clamp, l.conv_clamp)
        if i + 1 < len(layers):
            xs = h(x * Sn[i + 1][:, :, None, None])
    return x * net.output_scale

res = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cb = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
cm = int(sys.argv[3]) if len(sys.argv) > 3 else 128
net = sg3.make_synthesis("T", img_resolution=res, channel_base=cb, channel_max=cm)
torch.manual_seed(1)
ws = torch.randn(1, 16, 512)
ref = net(ws)
pix = lambda y: (y + 1) / 2
for cfg in [dict(filt_half=False, inter_half=True), dict(filt_half=True, inter_half=True), dict(filt_half=True, inter_half=True, dc=True)]:
    DC = cfg.pop('dc', False)
    out = emulate(net, ws, **cfg)
    cfg['dc'] = DC
    e = (pix(out).clamp(0, 1) - pix(ref).clamp(0, 1)).abs()
    print(cfg, "max-abs pix err %.3e  rms %.3e" % (e.max(), e.square().mean().sqrt()), flush=True)
