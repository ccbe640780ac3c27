"""Synthetic I420/RGBA content for benchmarks and tests (content A-D), seeded."""
import numpy as np

SEED = 20261018


def _blur(a, sigma=2.0):
    r = int(3 * sigma); k = np.exp(-0.5 * (np.arange(-r, r + 1) / sigma) ** 2); k /= k.sum()
    a = np.apply_along_axis(lambda v: np.convolve(np.pad(v, r, mode="wrap"), k, mode="valid"), 0, a)
    return np.apply_along_axis(lambda v: np.convolve(np.pad(v, r, mode="wrap"), k, mode="valid"), 1, a)


def _stretch(a, lo=16, hi=235):
    a = (a - a.min()) / max(a.max() - a.min(), 1e-9)
    return (lo + a * (hi - lo)).round().astype(np.uint8)


class Content:
    """kind: 'A' moving texture, 'B' screen-like, 'C' static, 'D' white noise, 'E' static scene with sensor noise (sigma 1.2) and one moving
    object -- the content background detection is for."""

    def __init__(self, kind, width, height, seed=SEED):
        self.kind, self.w, self.h = kind, width, height
        self.rng = np.random.default_rng(seed)
        w, h = width, height
        if kind in ("A", "C", "E"):
            self.ty = _stretch(_blur(self.rng.random((h + 64, w + 64))))
            self.tu = _stretch(_blur(self.rng.random((h // 2 + 32, w // 2 + 32))), 64, 192)
            self.tv = _stretch(_blur(self.rng.random((h // 2 + 32, w // 2 + 32))), 64, 192)
        elif kind == "B":
            y = np.full((h, w), 235, np.uint8)
            for _ in range(max(4, w * h // 40000)):
                x0, y0 = self.rng.integers(0, w - 16), self.rng.integers(0, h - 16)
                bw, bh = self.rng.integers(16, max(17, w // 3)), self.rng.integers(8, max(9, h // 6))
                y[y0:y0 + bh, x0:x0 + bw] = self.rng.integers(16, 235)
            txt = self.rng.random((h, w)) < 0.08
            mask = np.zeros((h, w), bool); mask[(np.arange(h) // 12 % 3 == 1)[:, None] & np.ones((1, w), bool)] = True
            y[txt & mask] = 16
            self.by = y
            self.bu = np.full((h // 2, w // 2), 128, np.uint8); self.bv = np.full((h // 2, w // 2), 128, np.uint8)

    def frame(self, t):
        w, h = self.w, self.h
        if self.kind == "A":
            tri = lambda v: v % 128 if v % 128 < 64 else 127 - v % 128   # bounce inside the 64-pixel margin: no jumps
            ox, oy = tri(3 * t), 63 - tri(2 * t)
            y = np.roll(self.ty, (oy, ox), (0, 1))[:h, :w].copy()
            u = np.roll(self.tu, (oy // 2, ox // 2), (0, 1))[:h // 2, :w // 2]
            v = np.roll(self.tv, (oy // 2, ox // 2), (0, 1))[:h // 2, :w // 2]
            for i, (vx, vy, val) in enumerate(((4, 0, 40), (-2, 2, 200), (0, -6, 120))):
                bw, bh = max(8, w // 8), max(8, h // 8)
                x0 = (w // 4 * (i + 1) + vx * t) % max(1, w - bw); y0 = (h // 4 * (i + 1) + vy * t) % max(1, h - bh)
                y[y0:y0 + bh, x0:x0 + bw] = val
        elif self.kind == "C":
            y, u, v = self.ty[:h, :w], self.tu[:h // 2, :w // 2], self.tv[:h // 2, :w // 2]
        elif self.kind == "E":
            r = np.random.default_rng(SEED + 104729 * t)
            noisy = lambda p, lo, hi: np.clip(p.astype(np.float32) + r.normal(0.0, 1.2, p.shape), lo, hi).round().astype(np.uint8)
            y = noisy(self.ty[:h, :w], 16, 235)
            u, v = noisy(self.tu[:h // 2, :w // 2], 16, 240), noisy(self.tv[:h // 2, :w // 2], 16, 240)
            bw, bh = max(16, w // 6), max(16, h // 6)
            x0, y0 = (w // 5 + 5 * t) % max(1, w - bw), (h // 3 + 2 * t) % max(1, h - bh)
            y[y0:y0 + bh, x0:x0 + bw] = 60
        elif self.kind == "B":
            # cumulative screen updates: frame t = base picture with the updates of steps 0..t applied (stateless for the caller)
            if getattr(self, "_b_t", None) is None or self._b_t > t:
                self._b_y, self._b_t = self.by.copy(), -1
            y = self._b_y
            mbw, mbh = (w + 15) // 16, (h + 15) // 16
            while self._b_t < t:
                self._b_t += 1
                r = np.random.default_rng(SEED + 7919 * self._b_t)
                for _ in range(max(1, mbw * mbh // 20)):
                    mx, my = r.integers(0, mbw), r.integers(0, mbh)
                    blk = y[my * 16:my * 16 + 16, mx * 16:mx * 16 + 16]
                    blk[...] = np.where(r.random(blk.shape) < 0.15, 16, r.integers(100, 235)).astype(np.uint8)
            u, v = self.bu, self.bv
        else:
            r = np.random.default_rng(SEED + t)
            y = r.integers(0, 256, (h, w), dtype=np.uint8)
            u = r.integers(0, 256, (h // 2, w // 2), dtype=np.uint8); v = r.integers(0, 256, (h // 2, w // 2), dtype=np.uint8)
        return np.concatenate([np.ascontiguousarray(y).ravel(), np.ascontiguousarray(u).ravel(), np.ascontiguousarray(v).ravel()])


def i420_to_rgba(i420, w, h):
    """Deterministic RGBA framebuffer made from an I420 frame (nearest chroma), for the RGBA input path."""
    y = i420[:w * h].reshape(h, w).astype(np.int32)
    u = i420[w * h:w * h * 5 // 4].reshape(h // 2, w // 2).astype(np.int32).repeat(2, 0).repeat(2, 1)
    v = i420[w * h * 5 // 4:].reshape(h // 2, w // 2).astype(np.int32).repeat(2, 0).repeat(2, 1)
    c, d, e = y - 16, u - 128, v - 128
    r = np.clip((298 * c + 409 * e + 128) >> 8, 0, 255); g = np.clip((298 * c - 100 * d - 208 * e + 128) >> 8, 0, 255)
    b = np.clip((298 * c + 516 * d + 128) >> 8, 0, 255)
    return np.stack([r, g, b, np.full_like(r, 255)], -1).astype(np.uint8)


def psnr(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)
