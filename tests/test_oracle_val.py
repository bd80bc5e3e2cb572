"""Validation path, CPU side: the oracle restatement (oracle/validation_oracle.py) and the
mirror's host logic (TrainDataset, color_to_class, centroid table) against goldens generated from the reference's own
CrossEntropy.forward / color_to_class / TrainDataset / step() by oracle/make_golden_val.py."""
import hashlib
import json

import numpy as np
import pytest
import torch

from oracle import validation_oracle as V
from tests._golden import GOLDEN

META = json.loads((GOLDEN / 'meta_val.json').read_text())


def _case(name):
    kw = dict(META['cases'][name])
    temperature = kw.pop('temperature')
    want_loss = kw.pop('loss')
    kw.pop('classes_present')
    feats, cls = V.synthetic_batch(**kw)
    g = np.load(GOLDEN / f'val_{name}.npz')
    assert abs(float(g['loss']) - want_loss) < 1e-12
    return feats, cls, temperature, float(g['loss']), torch.from_numpy(g['pred']).long()


@pytest.mark.parametrize('name', sorted(META['cases']))
def test_oracle_cross_entropy_matches_reference(name):
    feats, cls, temperature, want_loss, want_pred = _case(name)
    r, t, rc, tc = V.split_batch(feats, cls)
    loss, pred, prob = V.cross_entropy(r, t, rc, tc, 22, temperature)
    assert abs(float(loss) - want_loss) <= 1e-6 * want_loss
    assert torch.equal(pred, want_pred)
    assert torch.allclose(prob.sum(1), torch.ones_like(prob.sum(1)), atol=1e-5)
    assert max(META['cases'][name]['classes_present']) > 14      # the cases exercise the wide class range


def test_centroid_table_and_color_to_class():
    from src.utils.utils import annotation_centroids, color_to_class
    c = V.annotation_centroids()
    assert c.shape == (22, 3) and c.dtype == np.int32 and np.array_equal(c, annotation_centroids())
    assert c[9].tolist() == [191, 0, 0] and c[21].tolist() == [128, 64, 128]
    centroids = torch.Tensor(c).float()
    g = torch.Generator().manual_seed(5)
    img = centroids[torch.randint(0, 22, (2, 40, 56), generator=g)].permute(0, 3, 1, 2) \
        + torch.randint(-40, 41, (2, 3, 40, 56), generator=g).float()
    want = torch.from_numpy(np.load(GOLDEN / 'val_color_to_class.npz')['cls']).long()
    assert torch.equal(V.color_to_class(img, centroids), want)
    assert torch.equal(color_to_class(img, centroids), want)


def _sha(t):
    return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()


def test_train_dataset_and_step_match_reference(tmp_path):
    """The mirror's TrainDataset draws the same clips (bit-identical tensors) as the reference's under the same seed,
    and the oracle's validation step reproduces the reference's step() loss on them."""
    from src.utils.datasets import TrainDataset
    sc = META['step']
    root = V.write_synthetic_dataset(tmp_path, sc['n_videos'], sc['n_frames'], sc['H'], sc['W'], sc['seed'])
    ds = TrainDataset(root / 'JPEGImages/480p', root / 'Annotations/480p', frame_num=10, color_jitter=False)
    assert len(ds) == sc['dataset_len']
    loader = torch.utils.data.DataLoader(ds, batch_size=sc['bs'], shuffle=False, num_workers=0, drop_last=True)
    torch.manual_seed(sc['torch_seed'])
    batches = [(img, ann) for img, ann, _ in loader]
    assert len(batches) == sc['n_batches']
    assert [[_sha(i), _sha(a)] for i, a in batches] == sc['checksums']
    g = np.load(GOLDEN / 'val_step.npz')
    centroids = torch.Tensor(V.annotation_centroids()).float()
    low = [V.color_to_class(V.downsample_annotation(a.reshape(-1, 3, 256, 256)), centroids).numpy().astype(np.uint8)
           for _, a in batches[:3]]
    assert np.array_equal(np.stack(low), g['classes'][:3])
    loss, losses = V.validation_step(batches[:3], V.StubEmbedder(sc['stub_seed']), centroids)
    assert np.allclose(losses, g['batch_losses'][:3], rtol=1e-5)
    assert abs(float(g['batch_losses'].mean()) - float(g['loss'])) < 1e-6      # step() returns the mean of batch losses


def test_train_step_refuses_training_and_cpu():
    from src.model.loss import CrossEntropy
    from src.train import step
    with pytest.raises(NotImplementedError):
        step([], torch.nn.Identity(), CrossEntropy(), None, 0, torch.zeros(22, 3), 0, mode='train')
    if not torch.cuda.is_available():
        feats, cls = V.synthetic_batch(1, T=3, seed=1)
        r, t, rc, tc = V.split_batch(feats, cls)
        with pytest.raises(RuntimeError):          # no CPU fallback on the product path
            CrossEntropy()(r, t, rc, tc)
