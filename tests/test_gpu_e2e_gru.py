"""GPU end-to-end parity (north_star: <=1e-2 px end-point error after 32 GRU iterations with
identical random-init weights).

The reference model code cannot travel to the GPU box, so the loop here is a compact RAFT-Stereo
style update operator written for this test (motion encoder with the 1x1 conv on the 36 lookup
channels like core/update.py:68-84, a ConvGRU, a disparity head, x-only coordinate updates like
core/SMoEStereo.py:233-248).  It is run twice with the SAME weights and inputs: once on a plain
PyTorch fp32 restatement of the reference correlation ops (einsum / avg_pool2d / grid_sample, as
in core/corr.py:110-156) and once on this repo's CUDA path.  Test infrastructure only.
"""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


class TorchCorrBlock1D:
    """fp32 PyTorch restatement of the reference 'reg' block (checker, not product)."""

    def __init__(self, fmap1, fmap2, num_levels=4, radius=4):
        self.L, self.r = num_levels, radius
        B, D, H, W1 = fmap1.shape
        corr = torch.einsum("aijk,aijh->ajkh", fmap1, fmap2) / torch.sqrt(torch.tensor(float(D)))
        corr = corr.reshape(B * H * W1, 1, 1, -1)
        self.pyr = [corr]
        for _ in range(num_levels - 1):
            corr = F.avg_pool2d(corr, [1, 2], stride=[1, 2])
            self.pyr.append(corr)

    def __call__(self, coords):
        B, _, H, W = coords.shape
        x = coords[:, :1].permute(0, 2, 3, 1).reshape(B * H * W, 1, 1, 1)
        dx = torch.linspace(-self.r, self.r, 2 * self.r + 1, device=coords.device).view(-1, 1)
        outs = []
        for i, lv in enumerate(self.pyr):
            Wi = lv.shape[-1]
            x0 = dx + x / 2 ** i
            grid = torch.cat([2 * x0 / (Wi - 1) - 1, torch.zeros_like(x0)], dim=-1)
            outs.append(F.grid_sample(lv, grid, align_corners=True).view(B, H, W, -1))
        return torch.cat(outs, dim=-1).permute(0, 3, 1, 2).contiguous().float()


class MiniUpdate(nn.Module):
    def __init__(self, cor_planes=36, hidden=64):
        super().__init__()
        self.convc1 = nn.Conv2d(cor_planes, 64, 1)
        self.convc2 = nn.Conv2d(64, 64, 3, padding=1)
        self.convf1 = nn.Conv2d(1, 32, 7, padding=3)
        self.conv = nn.Conv2d(96, 63, 3, padding=1)
        self.convz = nn.Conv2d(hidden + 64, hidden, 3, padding=1)
        self.convr = nn.Conv2d(hidden + 64, hidden, 3, padding=1)
        self.convq = nn.Conv2d(hidden + 64, hidden, 3, padding=1)
        self.head = nn.Sequential(nn.Conv2d(hidden, 64, 3, padding=1), nn.ReLU(), nn.Conv2d(64, 1, 3, padding=1))

    def forward(self, net, corr, flow_x):
        c = F.relu(self.convc2(F.relu(self.convc1(corr))))
        f = F.relu(self.convf1(flow_x))
        m = torch.cat([F.relu(self.conv(torch.cat([c, f], 1))), flow_x], 1)
        hx = torch.cat([net, m], 1)
        z = torch.sigmoid(self.convz(hx))
        r = torch.sigmoid(self.convr(hx))
        q = torch.tanh(self.convq(torch.cat([r * net, m], 1)))
        net = (1 - z) * net + z * q
        return net, self.head(net)


def run_loop(block_cls, fmap1, fmap2, upd, net0, iters):
    B, _, H, W = fmap1.shape
    ys, xs = torch.meshgrid(torch.arange(H, device=fmap1.device), torch.arange(W, device=fmap1.device),
                            indexing="ij")
    coords0 = torch.stack([xs, ys], 0).float()[None].repeat(B, 1, 1, 1)
    coords1 = coords0.clone()
    corr_fn = block_cls(fmap1, fmap2, num_levels=4, radius=4)
    net = net0.clone()
    for _ in range(iters):
        coords1 = coords1.detach()
        corr = corr_fn(coords1)
        flow = coords1 - coords0
        net, delta = upd(net, corr, flow[:, :1])
        coords1 = coords1.clone()
        coords1[:, :1] = coords1[:, :1] + delta          # delta_flow[:,1] = 0
    return (coords1 - coords0)[:, 0]


@pytest.mark.parametrize("shape", [(1, 256, 40, 96), (2, 256, 24, 64)])
def test_epe_after_32_gru_iterations(shape):
    import smoe_stereo_b200 as sb
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device("cuda", 0)
    B, C, H, W = shape
    g = torch.Generator(device="cpu").manual_seed(1234)
    # correlated feature maps: fmap2 is fmap1 shifted by a smooth disparity plus noise, so the
    # volume has a ridge and the update operator moves the estimate (not a fixed point at 0)
    f1 = torch.randn(B, C, H, W, generator=g)
    shift = 6
    f2 = torch.roll(f1, shifts=-shift, dims=3) + 0.3 * torch.randn(B, C, H, W, generator=g)
    f1, f2 = f1.to(dev), f2.to(dev)
    torch.manual_seed(7)
    upd = MiniUpdate().to(dev).eval()
    net0 = torch.tanh(torch.randn(B, 64, H, W, generator=g)).to(dev)
    with torch.no_grad():
        ref = run_loop(TorchCorrBlock1D, f1, f2, upd, net0, 32)
        got = run_loop(sb.CorrBlock1D, f1, f2, upd, net0, 32)
        fast = run_loop(sb.CorrBlockFast1D, f1, f2, upd, net0, 32)
    moved = ref.abs().mean().item()
    epe = (ref - got).abs().mean().item()
    epe_max = (ref - got).abs().max().item()
    epe_fast = (ref - fast).abs().mean().item()
    print(f"[e2e {shape}] mean |disp| {moved:.3f} px, EPE {epe:.2e} px (max {epe_max:.2e}), "
          f"Fast1D EPE {epe_fast:.2e}")
    assert moved > 0.05, "the loop did not move the estimate; the test would be vacuous"
    assert epe <= 1e-2
    assert epe_fast <= 1e-2
