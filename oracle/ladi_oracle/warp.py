"""TEST INFRASTRUCTURE ONLY (see oracle/ladi_oracle/__init__.py) -- fp32 CPU restatement of the cloth-warping front-end
(SURVEY.md section 8(f) row 2), inference path only:

  * `ConvNet_TPS`   <- /root/reference/src/models/ConvNet_TPS.py:287-337 (FeatureExtraction :29-55, FeatureL2Norm :58-66,
                       FeatureCorrelation :69-81, FeatureRegression :91-123, TPSGridGen :143-193, BoundedGridLocNet :196-225).  The
                       six regularisation scalars that the reference forward also returns (:207-225) are training losses (and call
                       `.cuda()` unconditionally); the restatement returns (grid, control points) only.
  * `UNetVanilla`   <- /root/reference/src/models/UNet.py:4-34 + unet_parts.py (bilinear=True variant, hubconf.py:58).
  * `warp_cloth`    <- the batch body of /root/reference/src/inference.py:236-263 (antialiased resizes to 256x192, TPS grid, grid
                       up-sampling, grid_sample with border padding, refinement, clamp).

Sub-module names are the reference's, so its checkpoints' state-dict keys load unchanged.  Parity status: pinned against the reference
classes themselves (imported unmodified, `torch.Tensor.cuda` patched to a no-op on this CPU-only container) with shared weights in
tests/test_oracle_pins.py (runs where /root/reference exists) and through tests/golden/warp_small.npz.
"""
import itertools

import torch
import torch.nn as nn
import torch.nn.functional as F


class FeatureExtraction(nn.Module):
    """conv4x4/s2 -> ReLU -> BN, (n_layers more of the same, widths doubling up to 512), conv3x3 -> ReLU -> BN, conv3x3 -> ReLU."""

    def __init__(self, input_nc, ngf=64, n_layers=3):
        super().__init__()
        seq = [nn.Conv2d(input_nc, ngf, 4, 2, 1), nn.ReLU(True), nn.BatchNorm2d(ngf)]
        for i in range(n_layers):
            cin = min(2 ** i * ngf, 512)
            cout = 2 ** (i + 1) * ngf if 2 ** i * ngf < 512 else 512
            seq += [nn.Conv2d(cin, cout, 4, 2, 1), nn.ReLU(True), nn.BatchNorm2d(cout)]
        seq += [nn.Conv2d(512, 512, 3, 1, 1), nn.ReLU(True), nn.BatchNorm2d(512), nn.Conv2d(512, 512, 3, 1, 1), nn.ReLU(True)]
        self.model = nn.Sequential(*seq)

    def forward(self, x):
        return self.model(x)


class FeatureRegression(nn.Module):
    def __init__(self, input_nc=192, output_dim=6, flat=64 * 4 * 3):
        super().__init__()
        self.conv = nn.Sequential(
            nn.Conv2d(input_nc, 512, 4, 2, 1), nn.BatchNorm2d(512), nn.ReLU(True),
            nn.Conv2d(512, 256, 4, 2, 1), nn.BatchNorm2d(256), nn.ReLU(True),
            nn.Conv2d(256, 128, 3, padding=1), nn.BatchNorm2d(128), nn.ReLU(True),
            nn.Conv2d(128, 64, 3, padding=1), nn.BatchNorm2d(64), nn.ReLU(True))
        self.linear = nn.Linear(flat, output_dim)

    def forward(self, x):
        return torch.tanh(self.linear(self.conv(x).reshape(x.shape[0], -1)))


def control_points(r=0.9, grid=5):
    """5x5 lattice in [-0.9, 0.9]^2, row-major over (y, x), stored as (x, y)  (ConvNet_TPS.py:300-307)."""
    step = 2.0 * r / (grid - 1)
    axis = [-r + i * step for i in range(grid)]
    return torch.tensor([[x, y] for y, x in itertools.product(axis, axis)], dtype=torch.float32)


def radial(a, b):
    """phi(r) = r^2 log r = 0.5 d2 log d2 between point sets a [N,2], b [M,2]; 0 at coincident points (ConvNet_TPS.py:127-139)."""
    d = a[:, None, :] - b[None, :, :]
    d2 = d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]
    u = 0.5 * d2 * torch.log(d2)
    return torch.where(torch.isnan(u), torch.zeros_like(u), u)


def tps_matrices(height, width, ctrl):
    """-> (inverse_kernel [N+3, N+3], target_coordinate_repr [H*W, N+3]) of TPSGridGen.__init__ (ConvNet_TPS.py:145-181)."""
    n = ctrl.shape[0]
    k = torch.zeros(n + 3, n + 3)
    k[:n, :n] = radial(ctrl, ctrl)
    k[:n, n] = 1
    k[n, :n] = 1
    k[:n, n + 1:] = ctrl
    k[n + 1:, :n] = ctrl.t()
    ys, xs = torch.meshgrid(torch.arange(height, dtype=torch.float32), torch.arange(width, dtype=torch.float32), indexing="ij")
    coord = torch.stack([xs.reshape(-1) * 2 / (width - 1) - 1, ys.reshape(-1) * 2 / (height - 1) - 1], dim=1)
    rep = torch.cat([radial(coord, ctrl), torch.ones(height * width, 1), coord], dim=1)
    return torch.inverse(k), rep


class TPSGridGen(nn.Module):
    def __init__(self, height, width, ctrl):
        super().__init__()
        inv, rep = tps_matrices(height, width, ctrl)
        self.register_buffer("inverse_kernel", inv)
        self.register_buffer("padding_matrix", torch.zeros(3, 2))
        self.register_buffer("target_coordinate_repr", rep)

    def forward(self, pts):
        y = torch.cat([pts, self.padding_matrix.expand(pts.shape[0], 3, 2)], dim=1)
        return self.target_coordinate_repr @ (self.inverse_kernel @ y)


class BoundedGridLocNet(nn.Module):
    def __init__(self, ctrl, input_nc=192, flat=768):
        super().__init__()
        self.regression = FeatureRegression(input_nc, ctrl.numel(), flat)
        with torch.no_grad():  # ConvNet_TPS.py:199-203: start at the identity warp
            self.regression.linear.bias.copy_(torch.atanh(ctrl).view(-1))
            self.regression.linear.weight.zero_()

    def forward(self, x):
        return self.regression(x).view(x.shape[0], -1, 2)


class ConvNet_TPS(nn.Module):
    def __init__(self, height, width, input_nc=6, n_layer=4):
        super().__init__()
        self.height, self.width = height, width
        ctrl = control_points()
        self.extractionA = FeatureExtraction(3, 64, n_layer)
        self.extractionB = FeatureExtraction(input_nc, 64, n_layer)
        fh, fw = height // 2 ** (n_layer + 1), width // 2 ** (n_layer + 1)
        self.loc_net = BoundedGridLocNet(ctrl, fh * fw, 64 * (fh // 4) * (fw // 4))
        self.gridGen = TPSGridGen(height, width, ctrl)

    def forward(self, a, b):
        norm = lambda f: f / torch.sqrt((f * f).sum(1, keepdim=True) + 1e-6)
        fa, fb = norm(self.extractionA(a)), norm(self.extractionB(b))
        n, c, h, w = fa.shape
        corr = fb.view(n, c, h * w).transpose(1, 2) @ fa.transpose(2, 3).reshape(n, c, h * w)       # [n, iB, iA] with iA = x*h + y
        corr = corr.view(n, h, w, h * w).permute(0, 3, 1, 2)                                          # [n, iA, yB, xB]
        pts = self.loc_net(corr)
        return self.gridGen(pts).view(n, self.height, self.width, 2), pts


class DoubleConv(nn.Module):
    def __init__(self, cin, cout, mid=None):
        super().__init__()
        mid = mid or cout
        self.double_conv = nn.Sequential(nn.Conv2d(cin, mid, 3, padding=1, bias=False), nn.BatchNorm2d(mid), nn.ReLU(True),
                                         nn.Conv2d(mid, cout, 3, padding=1, bias=False), nn.BatchNorm2d(cout), nn.ReLU(True))

    def forward(self, x):
        return self.double_conv(x)


class Down(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.maxpool_conv = nn.Sequential(nn.MaxPool2d(2), DoubleConv(cin, cout))

    def forward(self, x):
        return self.maxpool_conv(x)


class Up(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.up = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True)
        self.conv = DoubleConv(cin, cout, cin // 2)

    def forward(self, x1, x2):
        x1 = self.up(x1)
        dy, dx = x2.shape[2] - x1.shape[2], x2.shape[3] - x1.shape[3]
        x1 = F.pad(x1, [dx // 2, dx - dx // 2, dy // 2, dy - dy // 2])
        return self.conv(torch.cat([x2, x1], dim=1))


class OutConv(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, 1)

    def forward(self, x):
        return self.conv(x)


class UNetVanilla(nn.Module):
    def __init__(self, n_channels, n_classes, bilinear=True, widths=(64, 128, 256, 512, 1024)):
        super().__init__()
        assert bilinear, "the reference hub constructor builds the bilinear variant (hubconf.py:58)"
        w = widths
        self.inc = DoubleConv(n_channels, w[0])
        self.down1, self.down2, self.down3 = Down(w[0], w[1]), Down(w[1], w[2]), Down(w[2], w[3])
        self.down4 = Down(w[3], w[4] // 2)
        self.up1, self.up2, self.up3 = Up(w[4], w[3] // 2), Up(w[3], w[2] // 2), Up(w[2], w[1] // 2)
        self.up4 = Up(w[1], w[0])
        self.outc = OutConv(w[0], n_classes)

    def forward(self, x):
        x1 = self.inc(x)
        x2 = self.down1(x1)
        x3 = self.down2(x2)
        x4 = self.down3(x3)
        x = self.up1(self.down4(x4), x4)
        x = self.up2(x, x3)
        x = self.up3(x, x2)
        return self.outc(self.up4(x, x1))


def _resize(x, size):
    return F.interpolate(x, size=size, mode="bilinear", antialias=True, align_corners=False)  # == torchvision resize(BILINEAR, antialias)


def warp_cloth(tps, refinement, cloth, im_mask, pose_map, low=(256, 192), return_all=False):
    """src/inference.py:236-263: -> warped_cloth NCHW fp32 in [-1, 1] (the `warped_cloth` argument of the try-on pipeline)."""
    size = tuple(cloth.shape[-2:])
    agnostic = torch.cat([_resize(im_mask.float(), low), _resize(pose_map.float(), low)], dim=1)
    low_grid, pts = tps(_resize(cloth.float(), low), agnostic)
    grid = _resize(low_grid.permute(0, 3, 1, 2), size).permute(0, 2, 3, 1)
    coarse = F.grid_sample(cloth.float(), grid, padding_mode="border", align_corners=False)
    out = refinement(torch.cat([im_mask.float(), pose_map.float(), coarse], dim=1)).clamp(-1, 1)
    return (out, coarse, low_grid, pts) if return_all else out
