"""Tolerances used by the parity tests (north_star: fp32 results within 1e-5 relative).

Box corners are differences of two pixel-scale terms (x1 = cx - w/2), so one ulp of sigmoid/exp in w
shows up as an *absolute* error proportional to the box size even when the corner itself is near 0.
The floor of the relative test is therefore max(|ref|, box size, 1 px)."""
import torch

RTOL = 1e-5


def close(a, b, scale=1.0, rtol=RTOL):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return bool(((a - b).abs() <= rtol * torch.clamp(b.abs(), min=scale)).all())


def boxes_close(a, b, rtol=RTOL):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    size = torch.maximum((b[..., 2] - b[..., 0]).abs(), (b[..., 3] - b[..., 1]).abs()).clamp(min=1.0)
    return bool(((a - b).abs() <= rtol * torch.maximum(b.abs(), size[..., None])).all())
