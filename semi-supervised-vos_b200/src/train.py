"""`step`: one pass of a loader through model + criterion, the reference's src/train.py:155-216 in 'val' mode.
Training itself (optimizer, back-propagation through the affinity, checkpoints, TensorBoard; train.py:25-152) is out
of scope: the engine computes forward values only."""
import numpy as np
import torch
import torch.nn.functional
from tqdm import tqdm

from src.config import Config
from src.utils.utils import color_to_class


def step(loader, model, criterion, optimizer, epoch, centroids, batches, mode='train'):
    if mode == 'train':
        raise NotImplementedError("step(mode='train') needs gradients of the propagation; only mode='val' is built")
    model = model.eval()
    criterion.num_classes = centroids.shape[0]
    losses = []
    for img, annotation, _ in tqdm(loader, desc=f'Validating epoch {epoch}.', total=batches):
        B, T, C, H, W = img.shape
        # annotations: RGB -> stride-8 nearest samples -> index of the nearest centroid (train.py:165-173)
        low = torch.nn.functional.interpolate(annotation.reshape(-1, 3, H, W).to(Config.DEVICE, non_blocking=True),
                                              scale_factor=Config.SCALE, mode='nearest')
        H_d, W_d = low.shape[-2:]
        classes = color_to_class(low, centroids).reshape(B, T, H_d, W_d)
        with torch.no_grad():
            feats = model(img.reshape(-1, C, H, W).to(Config.DEVICE, non_blocking=True))
        feats = feats.reshape(B, T, feats.shape[1], H_d, W_d)
        # first T-1 frames are the references, the last one the target (train.py:181-184); the class maps go to the
        # criterion as indices -- the one-hot tensor of train.py:206 would only be arg-maxed back
        loss = criterion(feats[:, :-1], feats[:, -1], classes[:, :-1], classes[:, -1], None, None, False)
        losses.append(loss)
    return float(torch.stack(losses).mean().item()) if losses else float(np.array([]).mean())
