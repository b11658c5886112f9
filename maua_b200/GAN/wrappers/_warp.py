"""2x3 pixel-space matrices of the feature-map warps StyleGAN2Synthesizer hooks onto layer outputs
(maua/GAN/wrappers/stylegan2.py:153-194 via kornia.geometry.transform translate / rotate / scale; kornia is an absent,
un-pinned dependency, so its published matrix construction is restated: kornia/geometry/transform/affwarp.py +
imgwarp.py get_rotation_matrix2d).  Tiny host arithmetic; the warp itself is sg2_warp_kernel (csrc/sg2.cu), which
wants the INVERSE matrices: destination pixel (x, y, 1) -> source pixel."""
import torch


def _eye(B):
    return torch.eye(3, dtype=torch.float64).unsqueeze(0).repeat(B, 1, 1)


def translation_matrix(translation):
    """kT.translate: translation [B,2] in pixels (x, y) -> [B,3,3]."""
    t = translation.detach().cpu().double().reshape(-1, 2)
    m = _eye(t.shape[0])
    m[:, 0, 2] = t[:, 0]
    m[:, 1, 2] = t[:, 1]
    return m


def rotation_scale_matrix(angle_deg, scale, center, h, w):
    """kT.rotate / kT.scale: shift(c) @ rot(angle) @ diag(scale) @ shift(-c); centre default ((w-1)/2, (h-1)/2);
    positive angles rotate anti-clockwise; a 1-d scale applies to both axes."""
    a = torch.deg2rad(torch.as_tensor(angle_deg).detach().cpu().double().reshape(-1))
    s = torch.as_tensor(scale).detach().cpu().double()
    s = s.reshape(-1, 2) if (s.ndim == 2 and s.shape[-1] == 2) else s.reshape(-1, 1).repeat(1, 2)
    B = max(a.shape[0], s.shape[0])
    a = a.repeat(B) if a.shape[0] == 1 and B > 1 else a
    s = s.repeat(B, 1) if s.shape[0] == 1 and B > 1 else s
    if center is None:
        c = torch.tensor([(w - 1) / 2, (h - 1) / 2], dtype=torch.float64).unsqueeze(0).repeat(B, 1)
    else:
        c = torch.as_tensor(center).detach().cpu().double().reshape(-1, 2)
        c = c.repeat(B, 1) if c.shape[0] == 1 and B > 1 else c
    shift, shift_inv, rot, scl = _eye(B), _eye(B), _eye(B), _eye(B)
    shift[:, :2, 2] = c
    shift_inv[:, :2, 2] = -c
    rot[:, 0, 0] = torch.cos(a); rot[:, 0, 1] = torch.sin(a)
    rot[:, 1, 0] = -torch.sin(a); rot[:, 1, 1] = torch.cos(a)
    scl[:, 0, 0] = s[:, 0]; scl[:, 1, 1] = s[:, 1]
    return shift @ rot @ scl @ shift_inv


def inverse_2x3(m):
    """[B,3,3] forward matrices -> float32 [B,2,3] destination -> source maps (what warp_affine samples with)."""
    return torch.linalg.inv(m)[:, :2, :].to(torch.float32).contiguous()
