"""Colour mapping of a normalised scalar (mirror of the reference's eng/colormap.py:8-53), vectorised over torch tensors.

A channel is a clamped tent: height h, centre c, half-width wl left of the centre and wr right of it,
    map(x) = clamp((w - |clamp(x) - c|) / w * h),   w = wl if x < c else wr            (colormap.py:20-27)
and the default map is the "jet" triple (colormap.py:36-38, 48-51)."""


class ColorMap:
    def __init__(self, h, wl, wr, c):
        self.h, self.wl, self.wr, self.c = float(h), float(wl), float(wr), float(c)

    @staticmethod
    def clamp(x):
        return x.clamp(0.0, 1.0)

    def map(self, x):
        import torch
        w = torch.where(x < self.c, torch.full_like(x, self.wl), torch.full_like(x, self.wr))
        return self.clamp((w - (self.clamp(x) - self.c).abs()) / w * self.h)


jetR, jetG, jetB = ColorMap(1.5, .37, .37, .75), ColorMap(1.5, .37, .37, .5), ColorMap(1.5, .37, .37, .25)
bwrR, bwrG, bwrB = ColorMap(1.0, .25, 1, .5), ColorMap(1.0, .5, .5, .5), ColorMap(1.0, 1, .25, .5)
coolwarmR, coolwarmG, coolwarmB = ColorMap(0.9, .25, 1, .5), ColorMap(0.9, .5, .5, .5), ColorMap(0.9, 1, .25, .5)


def color_map(c):
    """(n,) normalised values -> (n, 3) float32 RGB with the jet map (the reference's active choice)."""
    import torch
    c = c.to(torch.float32)
    return torch.stack([jetR.map(c), jetG.map(c), jetB.map(c)], dim=1)
