"""NumPy restatement of the north-star extensions of the correlation filter: Gaussian kernel correlation, sub-pixel
peak, padding, target-sized labels.  TEST INFRASTRUCTURE.  The reference implements none of these (its filter is the linear-kernel DCF,
trackers/kcf.cpp:269-428), so there is no reference oracle: this file IS the specification the CUDA path is checked against (parity
unpinned by the reference).  It follows Henriques et al., "High-Speed Tracking with Kernelized Correlation Filters" (train / detect /
gaussian_correlation), on the reference's fHOG features (taken from the oracle) with the reference's window, label shift, learning
rate, lambda and peak search (first maximum in memory order, wrap, box shift in float with truncation, kcf.cpp:402-428).
All transforms in float64 via numpy.fft; everything the CUDA kernel computes in float32 is compared with a tolerance."""
import numpy as np

CELL = 4


def pad_box(l, t, b, r, p):
    """csrc/mot_internal.h: kcf_pad_box (integer arithmetic, floor division)"""
    w, h = r - l + 1, b - t + 1
    pw, ph = int(np.float32(w) * np.float32(p)), int(np.float32(h) * np.float32(p))
    nl, nt = (l + r - pw + 1) >> 1, (t + b - ph + 1) >> 1
    return nl, nt, nt + ph - 1, nl + pw - 1


def unpad_box(l, t, b, r, tw, th):
    nl, nt = (l + r - tw + 2) >> 1, (t + b - th + 2) >> 1
    return nl, nt, nt + th - 1, nl + tw - 1


def hann(n):
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / (n - 1))).astype(np.float32)          # sigpack window.h:34-48


def label_1d(n, sigma):
    """exp(-x^2 / (2 sigma^2)) for x = -n/2 .. , circularly shifted so that x = 0 sits at index 0 (kcf.cpp:96-122, 78-94)"""
    sinv = np.float32(1.0 / (np.float64(np.float32(sigma)) ** 2))
    g = np.zeros(n, np.float32)
    x0 = -(n // 2)
    for i in range(n):
        x = x0 + i
        g[(i + x0) % n] = np.float32(np.exp(-0.5 * x * x * np.float64(sinv)))
    return g


class KcfExt:
    def __init__(self, rows, cols, gaussian=False, sigma=0.5, subpixel=False, osf=0.0, lam=1e-4, lr=0.05):
        self.rows, self.cols = rows, cols
        self.hr, self.wc = rows // CELL, cols // CELL
        self.gaussian, self.sigma, self.subpixel, self.osf, self.lam, self.lr = gaussian, sigma, subpixel, osf, lam, lr
        self.win = (hann(self.wc)[:, None] * hann(self.hr)[None, :]).astype(np.float32)            # [j][i] = wx[j] * wy[i]
        self.model = None; self.alphaf = None
        self.pos = None; self.scale_h = self.scale_v = np.float32(1.0)
        self.tw = self.th = 0

    def _yf(self, tw, th):
        sig = np.float32(0.7289) if self.osf <= 0 else np.float32(np.sqrt(np.float32(tw) * np.float32(th))) * np.float32(self.osf) * np.float32(1.0 / CELL)
        lab = (label_1d(self.wc, sig)[:, None] * label_1d(self.hr, sig)[None, :]).astype(np.float32)   # outer product of floats, rounded to float
        return np.fft.fft2(lab.astype(np.float64))

    def _xf(self, H):
        """H: fHOG of the window, (32, wc, hr); 31 channels x Hann window, 2-D transform per channel"""
        x = (H[:31] * self.win[None]).astype(np.float32)
        return np.fft.fft2(x.astype(np.float64), axes=(1, 2))

    def _gcorr(self, xf, yf):
        n = self.hr * self.wc
        xx = (np.abs(xf) ** 2).sum() / n
        yy = (np.abs(yf) ** 2).sum() / n
        xy = np.real(np.fft.ifft2((xf * np.conj(yf)).sum(0)))
        k = np.exp(-np.maximum(0.0, xx + yy - 2.0 * xy) / (self.sigma ** 2 * n * 31))
        return np.fft.fft2(k)

    def update(self, H, window, target):
        """window, target: (l, t, b, r)"""
        xf = self._xf(H)
        first = self.model is None
        f = 1.0 if first else self.lr
        yf = self._yf(target[3] - target[0] + 1, target[2] - target[1] + 1)
        if self.gaussian:
            kf = self._gcorr(xf, xf)
            a = yf / (kf + self.lam)
        else:
            kf = (np.abs(xf) ** 2).sum(0) / (self.wc * self.hr * 31)                                # kcf.cpp:269-304
            a = np.real(yf) / (kf + self.lam)                                                        # real alpha, kcf.cpp:364-378
        self.alphaf = a if first else (1 - f) * self.alphaf + f * a
        self.model = xf if first else (1 - f) * self.model + f * xf
        self.pos = window
        self.scale_h = np.float32(window[3] - window[0] + 1) / np.float32(self.cols)
        self.scale_v = np.float32(window[2] - window[1] + 1) / np.float32(self.rows)
        self.tw, self.th = target[3] - target[0] + 1, target[2] - target[1] + 1

    def predict(self, H):
        """returns (new window (l,t,b,r), (sdv, sdh), response[j][i], pre-truncation float positions)"""
        zf = self._xf(H)
        if self.gaussian:
            resp = np.real(np.fft.ifft2(self.alphaf * self._gcorr(zf, self.model)))
        else:
            resp = np.real(np.fft.ifft2((zf * np.conj(self.model)).sum(0) * self.alphaf / (self.wc * self.hr * 31)))
        besti = int(np.argmax(resp.ravel()))                       # first maximum, j outer / i inner
        j0, i0 = besti // self.hr, besti % self.hr
        sdv = sdh = 0.0
        if self.subpixel:
            c = resp[j0, i0]
            up, dn = resp[j0, (i0 - 1) % self.hr], resp[j0, (i0 + 1) % self.hr]
            lf, rt = resp[(j0 - 1) % self.wc, i0], resp[(j0 + 1) % self.wc, i0]
            d1, d2 = 2 * c - dn - up, 2 * c - rt - lf
            sdv = 0.5 * (dn - up) / d1 if d1 != 0 else 0.0
            sdh = 0.5 * (rt - lf) / d2 if d2 != 0 else 0.0
        vd, hd = i0 + 1, j0 + 1
        if vd > self.hr // 2: vd -= self.hr
        if hd > self.wc // 2: hd -= self.wc
        dv = CELL * ((vd - 1) + sdv) * float(self.scale_v)
        dh = CELL * ((hd - 1) + sdh) * float(self.scale_h)
        l, t, b, r = self.pos
        fl = (l + dh, t + dv, b + dv, r + dh)
        self.pos = (int(fl[0]), int(fl[1]), int(fl[2]), int(fl[3]))                 # truncation toward zero, kcf.cpp:423-426
        return self.pos, (sdv, sdh), resp, fl
