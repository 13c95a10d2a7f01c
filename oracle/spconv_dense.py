"""ORACLE tooling: dense-grid stand-in for the spconv classes the reference imports
(spconv.pytorch.{core.SparseConvTensor, conv.SubMConv3d, conv.SparseConv3d, modules.SparseSequential};
call sites: /root/reference/ldm/models/diffusion/morphable_diffusion.py:19,253-254 and network.py:3-4,74-161).

spconv ("spconv-cu113" in the reference's requirements.txt:18, version otherwise unpinned) is not installed and not
vendored, and the reference has no test that pins results at this boundary => PARITY UNPINNED here.  This file
restates the published spconv-2.x semantics so the rest of the reference can run unmodified on CPU:
  * weights are stored [O, kd, kh, kw, I];
  * SubMConv3d: output active set == input active set; values = dense correlation over active neighbours;
  * SparseConv3d(k,s,p): output active wherever the receptive field holds an active input;
  * dense modules in a SparseSequential (BatchNorm1d, ReLU) act on the active rows only;
  * .dense() returns [B, C, D, H, W] with zeros at inactive voxels;
  * duplicate coordinates: LOWEST row index wins (rule fixed by this build; spconv leaves it undefined).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class SparseConvTensor:
    def __init__(self, features, indices, spatial_shape, batch_size):
        D, H, W = [int(s) for s in spatial_shape]
        C = features.shape[1]
        idx = indices.long()
        lin = ((idx[:, 0] * D + idx[:, 1]) * H + idx[:, 2]) * W + idx[:, 3]
        grid = torch.zeros(C, batch_size * D * H * W, dtype=features.dtype)
        order = torch.arange(lin.shape[0] - 1, -1, -1)
        grid[:, lin[order]] = features.t()[:, order]
        occ = torch.zeros(batch_size * D * H * W, dtype=features.dtype)
        occ[lin] = 1
        self.grid = grid.view(C, batch_size, D, H, W).permute(1, 0, 2, 3, 4).contiguous()
        self.mask = occ.view(batch_size, 1, D, H, W)

    def dense(self):
        return self.grid


class _SpConv(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=True, indice_key=None):
        super().__init__()
        k = kernel_size
        self.weight = nn.Parameter(torch.randn(out_channels, k, k, k, in_channels) * 0.05)
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None
        self.stride, self.padding, self.k = stride, padding, k

    def dense_weight(self):
        return self.weight.permute(0, 4, 1, 2, 3)


class SubMConv3d(_SpConv):
    def forward(self, x):
        out = SparseConvTensor.__new__(SparseConvTensor)
        out.grid = F.conv3d(x.grid, self.dense_weight(), self.bias, padding=self.k // 2) * x.mask
        out.mask = x.mask
        return out


class SparseConv3d(_SpConv):
    def forward(self, x):
        out = SparseConvTensor.__new__(SparseConvTensor)
        out.mask = F.max_pool3d(x.mask, self.k, self.stride, self.padding)
        out.grid = F.conv3d(x.grid, self.dense_weight(), self.bias, stride=self.stride, padding=self.padding) * out.mask
        return out


class SparseSequential(nn.Sequential):
    def forward(self, x):
        for m in self:
            if isinstance(m, _SpConv):
                x = m(x)
            elif isinstance(m, nn.BatchNorm1d):
                B, C = x.grid.shape[:2]
                flat = x.grid.permute(0, 2, 3, 4, 1).reshape(-1, C)
                y = m(flat).view(B, *x.grid.shape[2:], C).permute(0, 4, 1, 2, 3)
                x.grid = y * x.mask
            else:
                x.grid = m(x.grid) * x.mask
        return x
