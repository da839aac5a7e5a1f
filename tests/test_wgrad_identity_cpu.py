"""The index identity behind the cross-stacked mode of csrc/conv_wgrad.cu, checked in plain torch on the CPU:

    dW[o][i][ky][kx] = sum_p dy[o][p] * x[i][p + (ky-1, kx-1)]                      (autograd of lreq.py:126-156, padding 1)
                     = sum_q dy[o][q - (0, kx-1)] * x[i][q + (ky-1, 0)]            (q = p + (0, kx-1))

with q running over the pixel TILES of the map and everything outside the map read as zero (what the TMA loads of the
shifted tiles return): A rows = (kx, o) from three dy tiles shifted by one column, B columns = (ky, i) from three x tiles
shifted by one row, one product per tile.  Also with tiles that overhang the map (ragged sizes)."""
import pytest
import torch
import torch.nn.functional as F


def _tile(t, y0, x0, th, tw):
    """[C, th, tw] window of t [C, H, W] at (y0, x0), zeros outside the map (TMA out-of-bounds fill)."""
    c, h, w = t.shape
    out = t.new_zeros((c, th, tw))
    ys, xs = max(y0, 0), max(x0, 0)
    ye, xe = min(y0 + th, h), min(x0 + tw, w)
    if ye > ys and xe > xs:
        out[:, ys - y0:ye - y0, xs - x0:xe - x0] = t[:, ys:ye, xs:xe]
    return out


@pytest.mark.parametrize("n,cin,cout,h,w,th,tw", [(2, 4, 3, 8, 16, 1, 16), (1, 3, 5, 7, 21, 2, 8), (2, 2, 2, 5, 5, 4, 16)])
def test_cross_stacked_weight_gradient_identity(n, cin, cout, h, w, th, tw):
    g = torch.Generator().manual_seed(h * w)
    x = torch.randn(n, cin, h, w, generator=g, dtype=torch.float64)
    dy = torch.randn(n, cout, h, w, generator=g, dtype=torch.float64)
    wt = torch.zeros(cout, cin, 3, 3, dtype=torch.float64, requires_grad=True)
    (ref,) = torch.autograd.grad(F.conv2d(x, wt, padding=1), wt, dy)
    acc = torch.zeros(3 * cout, 3 * cin, dtype=torch.float64)                 # rows (kx, o), columns (ky, i)
    for s in range(n):
        for y0 in range(0, h, th):
            for x0 in range(0, w, tw):
                a = torch.cat([_tile(dy[s], y0, x0 + 1 - kx, th, tw) for kx in range(3)]).reshape(3 * cout, -1)
                b = torch.cat([_tile(x[s], y0 + ky - 1, x0, th, tw) for ky in range(3)]).reshape(3 * cin, -1)
                acc += a @ b.t()
    got = acc.view(3, cout, 3, cin).permute(1, 3, 2, 0)                        # [o][i][ky][kx]
    assert torch.allclose(got, ref, atol=1e-12)
