"""Golden vectors for the validation path, produced by the UNMODIFIED reference on CPU:
``CrossEntropy.forward`` (src/model/loss.py:45-66), ``color_to_class`` (src/utils/utils.py:45-56), ``TrainDataset``
(src/utils/datasets.py:15-109) and ``step(..., mode='val')`` (src/train.py:155-216).

Run in the build container (needs /root/reference):  python oracle/make_golden_val.py
Inputs are regenerated from seeds by oracle/validation_oracle.py; only outputs (losses, arg-max maps, class maps,
checksums of the dataset tensors) are stored under tests/golden/."""
import hashlib
import json
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
from oracle import reference_harness as RH          # noqa: E402
from oracle import validation_oracle as V           # noqa: E402

GOLDEN = REPO / 'tests' / 'golden'

CE_CASES = {   # name: synthetic_batch kwargs + temperature
    'ce_fp32': dict(B=2, seed=101, half=False, temperature=1.0),
    'ce_f16': dict(B=3, seed=102, half=True, temperature=1.0),
    'ce_temp': dict(B=2, seed=103, half=True, temperature=0.5, feat_scale=0.45),
    'ce_few_refs': dict(B=2, seed=104, half=True, temperature=1.0, T=4),
    'ce_noisy': dict(B=2, seed=105, half=True, temperature=1.0, noise=8.0),           # bad model: high loss, tiny p(true)
    'ce_noisy_fp32': dict(B=1, seed=106, half=False, temperature=1.0, noise=8.0, feat_scale=0.45),
}
STEP_CASE = dict(n_videos=2, n_frames=12, H=288, W=352, seed=7, bs=2, torch_seed=1234, stub_seed=3)


def sha(t: torch.Tensor) -> str:
    return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()


def main():
    ref = RH.import_validation(RH.import_reference('cpu'))
    centroids_file = np.load(ref.root / 'annotation_centroids.npy')
    assert np.array_equal(centroids_file, V.annotation_centroids()) and centroids_file.dtype == V.annotation_centroids().dtype
    centroids = torch.Tensor(centroids_file).float()
    meta = {'centroids': 'validation_oracle.annotation_centroids() == annotation_centroids.npy', 'cases': {}}

    for name, kw in CE_CASES.items():
        kw = dict(kw)
        temperature = kw.pop('temperature')
        feats, cls = V.synthetic_batch(**kw)
        d = 22
        r, t, rc, tc = V.split_batch(feats, cls)
        onehot = torch.zeros(r.shape[0], r.shape[1], d, *cls.shape[-2:]).scatter_(2, rc.unsqueeze(2), 1)   # train.py:206
        crit = ref.loss.CrossEntropy(temperature=temperature)
        loss, pred = crit(r, t, onehot, tc, None, None, True)
        o_loss, o_pred, _ = V.cross_entropy(r, t, rc, tc, d, temperature)
        assert abs(float(o_loss) - float(loss)) <= 1e-5 * abs(float(loss)) and torch.equal(o_pred, pred), (name, float(o_loss), float(loss))
        np.savez_compressed(GOLDEN / f'val_{name}.npz', loss=np.float64(loss.item()), pred=pred.numpy().astype(np.uint8))
        meta['cases'][name] = dict(kw, temperature=temperature, loss=float(loss),
                                   classes_present=sorted(set(cls.unique().tolist())))
        print(name, float(loss), 'accuracy', float((pred == tc).float().mean()))

    # color_to_class on off-centroid colours (JPEG-like noise around the palette)
    g = torch.Generator().manual_seed(5)
    img = centroids[torch.randint(0, 22, (2, 40, 56), generator=g)].permute(0, 3, 1, 2) + torch.randint(-40, 41, (2, 3, 40, 56), generator=g).float()
    want = ref.utils.color_to_class(img, centroids)
    assert torch.equal(want, V.color_to_class(img, centroids))
    np.savez_compressed(GOLDEN / 'val_color_to_class.npz', cls=want.numpy().astype(np.uint8))

    # TrainDataset + step(): a synthetic DAVIS-shaped tree, the reference's loader order and RNG draws
    sc = STEP_CASE
    with tempfile.TemporaryDirectory() as tmp:
        root = V.write_synthetic_dataset(tmp, sc['n_videos'], sc['n_frames'], sc['H'], sc['W'], sc['seed'])
        ds = ref.datasets.TrainDataset(root / 'JPEGImages/480p', root / 'Annotations/480p', frame_num=10, color_jitter=False)
        loader = torch.utils.data.DataLoader(ds, batch_size=sc['bs'], shuffle=False, num_workers=0, drop_last=True)
        torch.manual_seed(sc['torch_seed'])
        batches = [(img, ann) for img, ann, _ in loader]
        sums = [[sha(img), sha(ann)] for img, ann in batches]
        model = V.StubEmbedder(sc['stub_seed'])      # (its default init draws from the global RNG: build it first)
        torch.manual_seed(sc['torch_seed'])
        crit = ref.loss.CrossEntropy(temperature=1.0)
        with torch.no_grad():
            loss = ref.train.step(loader, model, crit, None, 0, centroids, len(loader), mode='val')
        o_loss, o_losses = V.validation_step(batches, model, centroids)
        assert abs(o_loss - loss) <= 1e-6 * abs(loss), (o_loss, loss)
        low_cls = [V.color_to_class(V.downsample_annotation(a.reshape(-1, 3, 256, 256)), centroids).numpy().astype(np.uint8)
                   for _, a in batches]
        np.savez_compressed(GOLDEN / 'val_step.npz', loss=np.float64(loss), batch_losses=np.array(o_losses),
                            classes=np.stack(low_cls))
        meta['step'] = dict(sc, n_batches=len(batches), loss=float(loss), checksums=sums, dataset_len=len(ds))
        print('step', float(loss), len(batches), 'batches')
    (GOLDEN / 'meta_val.json').write_text(json.dumps(meta, indent=1))


if __name__ == '__main__':
    main()
