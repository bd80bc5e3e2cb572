"""DAVIS region (J) and boundary (F) measures, the reference's src/utils/metrics.py:11-183 (itself the DAVIS toolkit's
`db_eval_iou` / `db_eval_boundary`), on numpy + scipy: the reference needs scikit-image for one dilation with a disk.

Offline CPU scoring -- used here as an accuracy check of the propagated masks; nothing of
the propagation path depends on it."""
import numpy as np
from scipy import ndimage


def evaluate_segmentation(annotation, segmentation, void_pixels=None, threshold=0.008):
    return eval_j(annotation, segmentation, void_pixels), eval_f(annotation, segmentation, void_pixels, threshold)


def eval_j(annotation, segmentation, void_pixels=None):
    """Jaccard index of two binary maps (or stacks of maps: reduced over the last two axes); 1 where the union is empty."""
    if annotation.shape != segmentation.shape:
        raise AssertionError(f'Annotation({annotation.shape}) and segmentation:{segmentation.shape} dimensions do not match.')
    a, s = annotation.astype(bool), segmentation.astype(bool)
    if void_pixels is None:
        keep = np.ones_like(s)
    else:
        if annotation.shape != void_pixels.shape:
            raise AssertionError(f'Annotation({annotation.shape}) and void pixels:{void_pixels.shape} dimensions do not match.')
        keep = ~void_pixels.astype(bool)
    inters = np.sum(s & a & keep, axis=(-2, -1))
    union = np.sum((s | a) & keep, axis=(-2, -1))
    with np.errstate(divide='ignore', invalid='ignore'):
        j = inters / union
    if np.ndim(j) == 0:
        return 1 if np.isclose(union, 0) else j
    j[np.isclose(union, 0)] = 1
    return j


def eval_f(annotation, segmentation, void_pixels=None, bound_th=0.008):
    """Boundary F-measure of one pair of maps, or per frame of a (T,H,W) stack."""
    assert annotation.shape == segmentation.shape
    assert void_pixels is None or annotation.shape == void_pixels.shape
    if annotation.ndim == 2:
        return f_measure(segmentation, annotation, void_pixels, bound_th=bound_th)
    if annotation.ndim != 3:
        raise ValueError(f'db_eval_boundary does not support tensors with {annotation.ndim} dimensions')
    return np.array([f_measure(segmentation[t], annotation[t], None if void_pixels is None else void_pixels[t], bound_th=bound_th)
                     for t in range(annotation.shape[0])], dtype=np.float64)


def disk(radius):
    """Flat disk footprint, x^2 + y^2 <= r^2 on the integer grid -r..r (skimage.morphology.disk)."""
    r = np.arange(-radius, radius + 1)
    xx, yy = np.meshgrid(r, r)
    return (xx ** 2 + yy ** 2 <= radius ** 2).astype(np.uint8)


def f_measure(foreground_mask, gt_mask, void_pixels=None, bound_th=0.008):
    """Precision / recall of the one-pixel boundaries of the two masks, each matched against the other's boundary dilated by
    a disk of ceil(bound_th * image diagonal) pixels; F = harmonic mean."""
    assert np.atleast_3d(foreground_mask).shape[2] == 1
    keep = np.ones(foreground_mask.shape, bool) if void_pixels is None else ~void_pixels.astype(bool)
    bound_pix = bound_th if bound_th >= 1 else np.ceil(bound_th * np.linalg.norm(foreground_mask.shape))
    fg_boundary = _seg2bmap(foreground_mask * keep)
    gt_boundary = _seg2bmap(gt_mask * keep)
    footprint = disk(bound_pix).astype(bool)
    fg_dil = ndimage.binary_dilation(fg_boundary, structure=footprint)
    gt_dil = ndimage.binary_dilation(gt_boundary, structure=footprint)
    n_fg, n_gt = int(fg_boundary.sum()), int(gt_boundary.sum())
    if n_fg == 0 and n_gt == 0:
        precision, recall = 1, 1
    elif n_fg == 0:
        precision, recall = 1, 0
    elif n_gt == 0:
        precision, recall = 0, 1
    else:
        precision = np.sum(fg_boundary & gt_dil) / float(n_fg)
        recall = np.sum(gt_boundary & fg_dil) / float(n_gt)
    return 0 if precision + recall == 0 else 2 * precision * recall / (precision + recall)


def _seg2bmap(seg, width=None, height=None):
    """Binary boundary map, one pixel wide, offset half a pixel towards the origin: a pixel is a boundary pixel when it
    differs from its east, south or south-east neighbour (last row / column: only along the edge; the corner never)."""
    seg = np.asarray(seg).astype(bool)
    assert np.atleast_3d(seg).shape[2] == 1
    h, w = seg.shape[:2]
    width = w if width is None else width
    height = h if height is None else height
    if (width, height) != (w, h):
        raise NotImplementedError('boundary maps are computed at the segmentation\'s own size (the reference never resizes)')
    east = np.zeros_like(seg)
    south = np.zeros_like(seg)
    diag = np.zeros_like(seg)
    east[:, :-1], south[:-1, :], diag[:-1, :-1] = seg[:, 1:], seg[1:, :], seg[1:, 1:]
    b = (seg ^ east) | (seg ^ south) | (seg ^ diag)
    b[-1, :] = seg[-1, :] ^ east[-1, :]
    b[:, -1] = seg[:, -1] ^ south[:, -1]
    b[-1, -1] = False
    return b
