"""Posterior post-processing on the device: Gaussian KDE on a grid.

Replaces the only other O(samples x grid) computation of the reference,
``gaussian_kde(qparams[0, :]).pdf(np.linspace(lo, hi, 1000))`` in ``RSF.plot_dist``
(RSF.py:717-737), with SciPy's default Scott bandwidth ``n**(-1/5) * std(ddof=1)``.
Plotting itself stays out of scope.
"""
import numpy as np

from . import _lib

KDE_POINTS = 1000          # RSF.py:717


def scott_bandwidth(n: int, std_ddof1: float) -> float:
    """scipy.stats.gaussian_kde default: factor n^(-1/(d+4)) with d = 1, times the sample s.d."""
    return float(n) ** (-0.2) * float(std_ddof1)


def gaussian_kde_pdf(samples, grid=None, lo=None, hi=None, points=KDE_POINTS):
    """KDE of a 1-D sample set (NumPy array or CUDA tensor of any shape, flattened) on ``grid``
    (or ``linspace(lo, hi, points)``).  Returns ``(grid, pdf)`` as NumPy arrays."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    x = samples if hasattr(samples, "is_cuda") else torch.as_tensor(np.asarray(samples, dtype=np.float64))
    x = x.to("cuda" if not x.is_cuda else x.device, dtype=torch.float64).reshape(-1).contiguous()
    n = x.numel()
    if n < 2:
        raise ValueError("need at least two samples")
    std = float(x.std(unbiased=True).item())
    if not std > 0:
        raise np.linalg.LinAlgError("singular data: all samples are equal (gaussian_kde raises too)")
    if grid is None:
        lo = float(x.min().item()) if lo is None else lo
        hi = float(x.max().item()) if hi is None else hi
        grid = np.linspace(lo, hi, points)
    g = torch.as_tensor(np.asarray(grid, dtype=np.float64)).to(x.device).contiguous()
    pdf = torch.empty_like(g)
    with torch.cuda.device(x.device):
        _lib.check(lib.rsfm_kde_grid(_lib.ptr(x), n, _lib.ptr(g), g.numel(), scott_bandwidth(n, std), _lib.ptr(pdf),
                                     _lib.current_stream(torch, x.device)), "rsfm_kde_grid")
    return g.cpu().numpy(), pdf.cpu().numpy()
