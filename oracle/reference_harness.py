"""Import and drive the UNMODIFIED reference (hynekdav/semi-supervised-VOS) on CPU.

TEST INFRASTRUCTURE ONLY (see oracle/propagation_oracle.py).  Used by ``oracle/make_golden.py``
(build container, where ``/root/reference`` exists) and by ``bench.py --impl reference`` when a
copy of the reference is present (``/root/reference`` or ``baseline/_ref``).  Nothing here is
on the product path.

The reference needs three compatibility shims on a modern stack; they
are applied to the *environment*, never to the reference's sources:
  1. ``numpy.int``            (src/model/predict.py:85, src/utils/datasets.py:145)
  2. ``PIL.Image.ANTIALIAS``  (src/utils/datasets.py:146)
  3. ``model_zoo.load_url``   (src/model/backbone/resnet.py:194 -- pretrained download, no network)
  4. ``skimage.morphology``   (src/model/triplet_miners.py:15, src/utils/metrics.py:6-7 -- not installed here) -- see
                               _skimage_shim()
"""
from __future__ import annotations

import importlib
import os
import sys
from pathlib import Path
from typing import Optional

REPO = Path(__file__).resolve().parent.parent
CANDIDATES = [Path('/root/reference'), REPO / 'baseline' / '_ref']


def find_reference() -> Optional[Path]:
    for c in CANDIDATES:
        if (c / 'src' / 'model' / 'predict.py').is_file():
            return c
    return None


def import_reference(device: str = 'cpu'):
    """Returns a namespace with the reference modules; raises if the reference is absent."""
    root = find_reference()
    if root is None:
        raise FileNotFoundError('reference sources not found in ' + ', '.join(map(str, CANDIDATES)))
    import numpy as np
    import PIL.Image
    if not hasattr(np, 'int'):
        np.int = int  # shim 1
    if not hasattr(PIL.Image, 'ANTIALIAS'):
        PIL.Image.ANTIALIAS = PIL.Image.LANCZOS  # shim 2
    for name in list(sys.modules):
        if name == 'src' or name.startswith('src.'):
            raise RuntimeError(f'a different `src` package ({name}) is already imported; run the '
                               'reference in its own process')
    sys.path.insert(0, str(root))
    try:
        import torch
        resnet = importlib.import_module('src.model.backbone.resnet')
        resnet.model_zoo.load_url = lambda url, *a, **k: {}  # shim 3
        ns = type('Reference', (), {})()
        ns.root = root
        ns.config = importlib.import_module('src.config')
        ns.config.Config.DEVICE = torch.device(device)
        ns.predict = importlib.import_module('src.model.predict')
        ns.utils = importlib.import_module('src.utils.utils')
        ns.inference_utils = importlib.import_module('src.utils.inference_utils')
        ns.vos_net = importlib.import_module('src.model.vos_net')
        ns.resnet = resnet
    finally:
        sys.path.remove(str(root))
    return ns


def _skimage_shim():
    """Shim 4: scikit-image is not installed here.  The reference imports it for (a) `skeletonize` (skeleton triplet miners
    only -- never called by the goldens) and (b) one grey dilation with a disk footprint in the F-measure
    (src/utils/metrics.py:92-94).  (b) is restated with scipy: `disk(r)` = {x^2 + y^2 <= r^2} on -r..r and
    `dilation(img, fp)` = maximum over the footprint with the image border padded by the minimum (scikit-image's documented
    behaviour); goldens that depend on it say so."""
    import types
    if 'skimage' in sys.modules:
        return
    try:
        importlib.import_module('skimage.morphology')
        return
    except ImportError:
        pass
    import numpy as np
    from scipy import ndimage
    sk, mo = types.ModuleType('skimage'), types.ModuleType('skimage.morphology')

    def skeletonize(*a, **k):
        raise RuntimeError('skimage is not installed: the skeleton miners cannot run in this container')

    def disk(radius, dtype=np.uint8):
        r = np.arange(-radius, radius + 1)
        xx, yy = np.meshgrid(r, r)
        return np.array((xx ** 2 + yy ** 2) <= radius ** 2, dtype=dtype)

    def dilation(image, footprint=None, out=None):
        return ndimage.grey_dilation(image, footprint=np.asarray(footprint, dtype=bool), mode='constant', cval=0)

    mo.skeletonize, mo.disk, mo.dilation = skeletonize, disk, dilation
    sk.morphology = mo
    sys.modules['skimage'], sys.modules['skimage.morphology'] = sk, mo


def import_validation(ns):
    """Adds the reference's loss / train / datasets modules (validation path) to `ns`."""
    _skimage_shim()
    sys.path.insert(0, str(ns.root))
    try:
        ns.loss = importlib.import_module('src.model.loss')
        ns.train = importlib.import_module('src.train')
        ns.datasets = importlib.import_module('src.utils.datasets')
    finally:
        sys.path.remove(str(ns.root))
    return ns


def import_evaluation(ns):
    """Adds the reference's metrics / evaluation modules to `ns`; numpy >= 2 accepts `np.bool` again."""
    _skimage_shim()
    sys.path.insert(0, str(ns.root))
    try:
        ns.metrics = importlib.import_module('src.utils.metrics')
        ns.evaluation = importlib.import_module('src.evaluation')
    finally:
        sys.path.remove(str(ns.root))
    return ns


def run_inference_single(ref, features, first_label_full, palette, workdir, video='clip',
                         sigma_1=8.0, sigma_2=21.0, frame_range=40, ref_num=9, temperature=1.0,
                         probability_propagation=False):
    """Drive the reference's REAL ``inference_single`` (src/utils/inference_utils.py:23-87) with a
    table-lookup 'model' (features precomputed) and record every ``predict`` result.

    Returns (masks (T-1,H,W) uint8 read back from the PNGs it wrote, list of (d,P) predictions).
    """
    import numpy as np
    import torch
    from PIL import Image

    T = features.shape[0]
    H, W = first_label_full.shape
    workdir = Path(workdir)
    ann_dir = workdir / 'Annotations' / '480p'
    (ann_dir / video).mkdir(parents=True, exist_ok=True)
    img = Image.fromarray(first_label_full.astype(np.uint8), mode='P')
    img.putpalette(palette)
    img.save(ann_dir / video / '00000.png')
    save = workdir / 'out'

    class TableModel:
        def __call__(self, inp):
            return features[int(inp[0, 0, 0, 0].item())][None]

    loader = [(torch.full((1, 1, H, W), float(t)), (video,)) for t in range(T)]
    recorded = []
    iu = ref.inference_utils
    real_predict = ref.predict.predict

    def recording_predict(*a, **k):
        out = real_predict(*a, **k)
        recorded.append(out.clone())
        return out

    iu.predict = recording_predict
    try:
        with torch.no_grad():
            iu.inference_single(TableModel(), loader, T, ann_dir, video, str(save), sigma_1,
                                sigma_2, frame_range, ref_num, temperature,
                                probability_propagation, True)
    finally:
        iu.predict = real_predict
    masks = np.stack([np.asarray(Image.open(save / video / f'{t:05d}.png')) for t in range(1, T)])
    return masks.astype(np.uint8), recorded


def run_inference_two_streams(ref, strategy, feats_a, feats_b, first_label_full, palette, workdir, size_b=None,
                              video='clip', sigma_1=8.0, sigma_2=21.0, frame_range=40, ref_num=9, temperature=1.0,
                              probability_propagation=False, reduction='mean', scale=1.15):
    """Drive the reference's REAL test-time-augmentation loops (src/utils/inference_utils.py:90-511:
    inference_hor_flip / inference_ver_flip / inference_2_scale / inference_multimodel) with table-lookup models.
    feats_a / feats_b: (T,K,h,w) embeddings of the two streams; size_b: (H,W) of the second input when it differs
    (2-scale).  Returns the fused masks (T-1,H,W) uint8 read back from the PNGs it wrote."""
    import numpy as np
    import torch
    from PIL import Image

    T = feats_a.shape[0]
    H, W = first_label_full.shape
    workdir = Path(workdir)
    ann_dir = workdir / 'Annotations' / '480p'
    (ann_dir / video).mkdir(parents=True, exist_ok=True)
    img = Image.fromarray(first_label_full.astype(np.uint8), mode='P')
    img.putpalette(palette)
    img.save(ann_dir / video / '00000.png')
    save = workdir / 'out'
    Hb, Wb = (H, W) if size_b is None else size_b

    def table(feats):
        return lambda inp: feats[int(inp[0, 0, 0, 0].item())][None]

    iu = ref.inference_utils
    common = (T, ann_dir, video, str(save), sigma_1, sigma_2, frame_range, ref_num, temperature, probability_propagation)
    with torch.no_grad():
        if strategy == 'multimodel':
            loader = [(torch.full((1, 1, H, W), float(t)), (video,)) for t in range(T)]
            iu.inference_multimodel(table(feats_a), table(feats_b), loader, *common, reduction, True)
        else:
            loader = [([torch.full((1, 1, H, W), float(t)), torch.full((1, 1, Hb, Wb), float(t))], (video,)) for t in range(T)]
            model = lambda inp: (feats_a if tuple(inp.shape[2:]) == (H, W) and not getattr(model, 'second', False) else feats_b)[int(inp[0, 0, 0, 0].item())][None]  # noqa: E731
            if (Hb, Wb) == (H, W):
                # same-sized inputs: the loops call model(input_l) then model(input_r) -- alternate
                state = {'n': 0}

                def model(inp):  # noqa: F811
                    state['n'] += 1
                    return (feats_a if state['n'] % 2 == 1 else feats_b)[int(inp[0, 0, 0, 0].item())][None]
            if strategy == 'hor-flip':
                iu.inference_hor_flip(model, loader, *common, reduction, True)
            elif strategy == 'vert-flip':
                iu.inference_ver_flip(model, loader, *common, reduction, True)
            elif strategy in ('2-scale', 'hor-2-scale'):
                iu.inference_2_scale(model, loader, *common, scale, reduction, strategy == 'hor-2-scale', True)
            else:
                raise ValueError(strategy)
    masks = np.stack([np.asarray(Image.open(save / video / f'{t:05d}.png')) for t in range(1, T)])
    return masks.astype(np.uint8)


def run_inference_3_scale(ref, feats_by_video, firsts, palette, workdir, size, scale=1.15, sigma_1=8.0, sigma_2=21.0,
                          frame_range=40, ref_num=9, temperature=1.0, probability_propagation=False):
    """Drive the reference's REAL ``inference_3_scale`` (src/utils/inference_utils.py:514-595) with a table-lookup
    model.  feats_by_video: {video: [feats at 0.9, feats at 1.0, feats at `scale`]}, each (T,K,h,w); firsts: {video:
    first annotation (H,W)}; size = (H,W) of the loader's frames.  The model is keyed by the (nearest-resized) input's
    shape and by the frame / video number encoded in the input.  Returns {video: (T-1, 480, 910) uint8 masks read back
    from the PNGs} -- the output size is hard-coded in the reference (:574)."""
    import numpy as np
    import torch
    from PIL import Image

    H, W = size
    workdir = Path(workdir)
    ann_dir = workdir / 'Annotations' / '480p'
    videos = sorted(feats_by_video)
    for v in videos:
        (ann_dir / v).mkdir(parents=True, exist_ok=True)
        img = Image.fromarray(firsts[v].astype(np.uint8), mode='P')
        img.putpalette(palette)
        img.save(ann_dir / v / '00000.png')
    save = workdir / 'out'
    by_shape = {}
    for k, s in enumerate((0.9, 1.0, scale)):
        by_shape[(int(np.ceil(H * s)), int(np.ceil(W * s)))] = k
    assert len(by_shape) == 3

    def model(inp):
        k = by_shape[tuple(inp.shape[2:])]
        v, t = int(inp[0, 0, 0, 0].item()) // 1000, int(inp[0, 0, 0, 0].item()) % 1000
        return feats_by_video[videos[v]][k][t][None]

    loader = [(torch.full((1, 1, H, W), float(1000 * vi + t)), (v,)) for vi, v in enumerate(videos)
              for t in range(feats_by_video[v][0].shape[0])]
    iu = ref.inference_utils
    with torch.no_grad():
        iu.inference_3_scale(model, loader, len(loader), ann_dir, videos[0], str(save), sigma_1, sigma_2, frame_range,
                             ref_num, temperature, probability_propagation, scale, True)
    out = {}
    for v in videos:
        T = feats_by_video[v][0].shape[0]
        out[v] = np.stack([np.asarray(Image.open(save / v / f'{t:05d}.png')) for t in range(1, T)]).astype(np.uint8)
    return out


def default_palette():
    pal = [0, 0, 0, 128, 0, 0, 0, 128, 0, 128, 128, 0, 0, 0, 128, 128, 0, 128, 0, 128, 128,
           128, 128, 128, 64, 0, 0, 192, 0, 0, 64, 128, 0, 192, 128, 0]
    return pal + [0] * (768 - len(pal))
