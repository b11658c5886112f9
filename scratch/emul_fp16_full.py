import torch, numpy as np, sys, time
import torch.nn.functional as F
from oracle import sg3
torch.set_num_threads(8)
def h(x): return x.half().float()
def emulate(net, ws, y_half=True, x_half=True, w_half=True, n32=0):
    ws_ = ws.float().unbind(1)
    x = net.input(ws_[0])
    names = net.layer_names
    layers = [getattr(net, n) for n in names]
    # styles
    def styles(l, w):
        s = l.affine(w)
        if l.is_torgb: s = s * (1/np.sqrt(l.in_channels*l.conv_kernel**2))
        return s
    S = [styles(l, w) for l, w in zip(layers, ws_[1:])]
    Sn = []
    for l, s in zip(layers, S):
        if not l.is_torgb: s = s * s.square().mean(1, keepdim=True).rsqrt()
        Sn.append(s)
    xs = x * Sn[0][:, :, None, None]
    if x_half: xs = h(xs)
    for i, l in enumerate(layers):
        W = l.weight
        y_half_, x_half_, w_half_ = (False, False, False) if i < n32 else (y_half, x_half, w_half)
        if not l.is_torgb:
            Wn = W * W.square().mean([1,2,3], keepdim=True).rsqrt()
            wsq = Wn.square().sum([2,3])  # [O,I]
            d = (Sn[i].square() @ wsq.t() + 1e-8).rsqrt()  # [B,O]
        else:
            Wn = W; d = torch.ones(ws.shape[0], l.out_channels)
        Wh = h(Wn) if w_half_ else Wn
        y = F.conv2d(xs, Wh, padding=l.conv_kernel-1) * d[:, :, None, None]
        if y_half_: y = h(y)
        gain = 1.0 if l.is_torgb else float(np.sqrt(2)); slope = 1.0 if l.is_torgb else 0.2
        x = sg3.filtered_lrelu_ref(y, fu=l.up_filter, fd=l.down_filter, b=l.bias, up=l.up_factor, down=l.down_factor,
                                   padding=l.padding, gain=gain, slope=slope, clamp=l.conv_clamp)
        if i + 1 < len(layers):
            xs = x * Sn[i+1][:, :, None, None]
            if x_half and i + 1 >= n32: xs = h(xs)
    return x * net.output_scale
res = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cb = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
cm = int(sys.argv[3]) if len(sys.argv) > 3 else 128
net = sg3.make_synthesis("T", img_resolution=res, channel_base=cb, channel_max=cm)
torch.manual_seed(1)
ws = torch.randn(1, 16, 512)
ref = net(ws)
pix = lambda y: (y + 1) / 2
for cfg in [dict(y_half=True, x_half=True, w_half=True), dict(n32=5)]:
    out = emulate(net, ws, **cfg)
    e = (pix(out).clamp(0,1) - pix(ref).clamp(0,1)).abs()
    print(cfg, "max-abs pix err %.3e  rms %.3e  | raw std %.3f" % (e.max(), e.square().mean().sqrt(), ref.std()))
