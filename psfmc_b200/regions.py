"""
ds9 region files -> boolean fitting-region masks, numpy only.

The reference turns a ds9 region file into a bad-pixel mask with
``~pyregion.open(f).as_imagecoord(hdr).get_filter().mask(shape)``
(/root/reference/psfMC/utils.py:92-95). pyregion is not a dependency of this
package; the shapes that matter for fitting-region masks (circle, ellipse, box,
each optionally excluded with a leading ``-``) are evaluated here for region
files in the ``image`` (or ``physical``) coordinate frame.

Conventions followed (pyregion 2.x ``as_region_filter`` / ``region_filter``):
* ds9 image coordinates are 1-based: pixel (row i, col j) has centre (j+1, i+1);
* a pixel is inside a circle when ``dx**2 + dy**2 <= r**2``;
* shapes are combined in file order: an included shape is OR-ed into the list,
  an excluded shape replaces the list by ``(OR of list so far) & ~shape``.
Sky-coordinate (fk5/icrs/galactic) region files need a WCS and are rejected
with a clear error rather than mis-evaluated.
"""
import re

import numpy as np

_SHAPE_RE = re.compile(r'^\s*([+-]?)\s*(\w+)\s*\(([^)]*)\)')
_FRAMES = ('image', 'physical', 'fk5', 'fk4', 'icrs', 'galactic', 'ecliptic',
           'j2000', 'b1950', 'wcs', 'linear', 'amplifier', 'detector')


class RegionFormatError(ValueError):
    pass


def parse_ds9(text):
    """Return a list of ``(name, exclude, [floats])`` in file order."""
    shapes = []
    frame = 'image'
    for raw_line in text.splitlines():
        for line in raw_line.split(';'):
            line = line.split('#')[0].strip()
            if not line or line.startswith('global'):
                continue
            lowered = line.lower()
            if lowered in _FRAMES:
                frame = lowered
                continue
            match = _SHAPE_RE.match(line)
            if not match:
                continue
            if frame not in ('image', 'physical'):
                raise RegionFormatError(
                    'region frame "{}" needs a WCS; only image-frame ds9 '
                    'regions are supported'.format(frame))
            sign, name, args = match.groups()
            try:
                coords = [float(tok.strip().rstrip('"\'di'))
                          for tok in re.split(r'[,\s]+', args.strip()) if tok]
            except ValueError:
                raise RegionFormatError('cannot parse region line: ' + line)
            shapes.append((name.lower(), sign == '-', coords))
    return shapes


def _inside(name, coords, xx, yy):
    if name == 'circle':
        xc, yc, rad = coords[:3]
        return (xx - (xc - 1)) ** 2 + (yy - (yc - 1)) ** 2 <= rad ** 2
    if name in ('ellipse', 'box', 'rotbox'):
        xc, yc, size_a, size_b = coords[:4]
        rot = np.deg2rad(coords[4]) if len(coords) > 4 else 0.0
        dx, dy = xx - (xc - 1), yy - (yc - 1)
        cos_r, sin_r = np.cos(rot), np.sin(rot)
        along = dx * cos_r + dy * sin_r
        across = -dx * sin_r + dy * cos_r
        if name == 'ellipse':
            return (along / size_a) ** 2 + (across / size_b) ** 2 <= 1.0
        return (np.abs(along) <= 0.5 * size_a) & (np.abs(across) <= 0.5 * size_b)
    raise RegionFormatError('unsupported ds9 shape: ' + name)


def region_mask(text, shape):
    """Boolean array, True where the region filter selects the pixel."""
    shapes = parse_ds9(text)
    if not shapes:
        raise RegionFormatError('no shapes found in region file')
    yy, xx = np.mgrid[0:shape[0], 0:shape[1]].astype(np.float64)
    selected = np.zeros(shape, dtype=bool)
    for name, exclude, coords in shapes:
        inside = _inside(name, coords, xx, yy)
        if exclude:
            selected &= ~inside
        else:
            selected |= inside
    return selected


def region_mask_from_file(filename, shape):
    with open(filename, 'r') as fobj:
        text = fobj.read()
    return region_mask(text, shape)
