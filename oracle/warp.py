"""CPU restatement of the kornia warps the reference's StyleGAN2 wrapper hooks onto feature maps
(maua/GAN/wrappers/stylegan2.py:153-194: kT.translate / kT.rotate / kT.scale, padding_mode="reflection").

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED: kornia is an absent third-party dependency the
reference does not pin (setup.py:59 `kornia`, no version) and the reference holds no test or golden vector for these
hooks.  What follows is kornia's published algorithm (kornia/geometry/transform/affwarp.py and imgwarp.py, 0.6 line):

  translate(x, t)            M = [[1, 0, tx], [0, 1, ty]]
  rotate(x, angle, center)   M = T(c) @ [[cos a, sin a], [-sin a, cos a]] @ T(-c), a in degrees, centre default ((W-1)/2, (H-1)/2)
  scale(x, s, center)        M = T(c) @ diag(sx, sy) @ T(-c)
  affine(x, M)               warp_affine(x, M, (H, W), "bilinear", padding_mode, align_corners=True)
  warp_affine                M_norm = N_dst @ M3x3 @ N_src^-1 with N = pixel -> [-1, 1] (2 / (size - 1), offset -1);
                             grid = F.affine_grid(inv(M_norm)[:, :2], align_corners=True);
                             F.grid_sample(x, grid, bilinear, padding_mode, align_corners=True)
"""
import torch
import torch.nn.functional as F


def _eye(B):
    return torch.eye(3, dtype=torch.float64).unsqueeze(0).repeat(B, 1, 1)


def translation_matrix(translation):
    """translation [B,2] (pixels: x, y) -> [B,3,3] float64."""
    t = translation.detach().cpu().double().reshape(-1, 2)
    m = _eye(t.shape[0])
    m[:, 0, 2] = t[:, 0]
    m[:, 1, 2] = t[:, 1]
    return m


def _center(center, B, h, w):
    if center is None:
        c = torch.tensor([(w - 1) / 2, (h - 1) / 2], dtype=torch.float64).unsqueeze(0).repeat(B, 1)
    else:
        c = torch.as_tensor(center).detach().cpu().double().reshape(-1, 2)
        if c.shape[0] == 1:
            c = c.repeat(B, 1)
    return c


def rotation_scale_matrix(angle_deg, scale, center, h, w):
    """get_rotation_matrix2d(center, angle, scale) as a [B,3,3] float64 matrix."""
    a = torch.deg2rad(torch.as_tensor(angle_deg).detach().cpu().double().reshape(-1))
    B = a.shape[0]
    s = torch.as_tensor(scale).detach().cpu().double()
    s = s.reshape(-1, 1).repeat(1, 2) if s.ndim < 2 or s.shape[-1] != 2 else s.reshape(-1, 2)
    if s.shape[0] == 1 and B > 1:
        s = s.repeat(B, 1)
    if B == 1 and s.shape[0] > 1:
        a = a.repeat(s.shape[0])
        B = s.shape[0]
    c = _center(center, B, h, w)
    shift, shift_inv, rot, scl = _eye(B), _eye(B), _eye(B), _eye(B)
    shift[:, :2, 2] = c
    shift_inv[:, :2, 2] = -c
    rot[:, 0, 0] = torch.cos(a); rot[:, 0, 1] = torch.sin(a)
    rot[:, 1, 0] = -torch.sin(a); rot[:, 1, 1] = torch.cos(a)
    scl[:, 0, 0] = s[:, 0]; scl[:, 1, 1] = s[:, 1]
    return shift @ rot @ scl @ shift_inv


def warp_affine(x, M, padding_mode="reflection"):
    """kornia warp_affine(x, M[:, :2], (H, W), 'bilinear', padding_mode, align_corners=True); M [B,3,3] float64."""
    B, _, H, W = x.shape
    M = M.expand(B, 3, 3) if M.shape[0] == 1 else M
    norm = torch.tensor([[2.0 / (W - 1), 0, -1], [0, 2.0 / (H - 1), -1], [0, 0, 1]], dtype=torch.float64)
    m_norm = norm @ M @ torch.linalg.inv(norm)
    theta = torch.linalg.inv(m_norm)[:, :2].to(x.dtype)
    grid = F.affine_grid(theta, [B, x.shape[1], H, W], align_corners=True)
    return F.grid_sample(x, grid, mode="bilinear", padding_mode=padding_mode, align_corners=True)


def translate(x, translation):
    return warp_affine(x, translation_matrix(translation))


def rotate(x, angle, center=None):
    return warp_affine(x, rotation_scale_matrix(angle, torch.ones(1), center, x.shape[-2], x.shape[-1]))


def scale(x, factor, center=None):
    f = torch.as_tensor(factor)
    return warp_affine(x, rotation_scale_matrix(torch.zeros(max(f.reshape(-1).shape[0] if f.ndim < 2 else f.shape[0], 1)), f, center,
                                                x.shape[-2], x.shape[-1]))
