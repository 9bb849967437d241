"""SSIM as the reference's drift gate computes it (tests/_ssim.py: Wang et al. 2004, 11 x 11 Gaussian window, sigma 1.5, K = (0.01, 0.03),
zero-padded 'constant' borders, mean over the map, per channel then averaged); restated for the wavefront golden pin."""
import numpy as np
from scipy.ndimage import convolve


def ssim(img1, img2, data_range=255.0, k1=0.01, k2=0.03, win_size=11, sigma=1.5) -> float:
    a, b = np.asarray(img1, np.float64), np.asarray(img2, np.float64)
    assert a.shape == b.shape
    if a.ndim == 3:
        return float(np.mean([ssim(a[..., c], b[..., c], data_range, k1, k2, win_size, sigma) for c in range(a.shape[2])]))
    x = np.arange(win_size, dtype=np.float64) - (win_size - 1) / 2.0
    g = np.exp(-0.5 * (x / sigma) ** 2)
    g /= g.sum()
    win = np.outer(g, g)
    f = lambda im: convolve(im, win, mode="constant", cval=0.0)
    c1, c2 = (k1 * data_range) ** 2, (k2 * data_range) ** 2
    mu1, mu2 = f(a), f(b)
    s1, s2, s12 = f(a * a) - mu1 * mu1, f(b * b) - mu2 * mu2, f(a * b) - mu1 * mu2
    return float(np.mean(((2 * mu1 * mu2 + c1) * (2 * s12 + c2)) / ((mu1 * mu1 + mu2 * mu2 + c1) * (s1 + s2 + c2))))
