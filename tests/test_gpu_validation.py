"""Validation path on the GPU: the engine-backed CrossEntropy / step() / `validation` command
against the reference goldens (tests/golden/val_*.npz, from the reference's own CrossEntropy.forward and step()) and
the oracle restatement (oracle/validation_oracle.py).  Also the 15..24-class capability of the index-label kernel that
the path needs (d = 22 annotation centroids), with and without the spatial prior.

Bars: loss within 1e-4 relative of the reference's, probabilities within 1e-3, arg-max maps >= 99.9 % equal."""
import json

import numpy as np
import pytest
import torch

from oracle import propagation_oracle as O
from oracle import validation_oracle as V
from oracle.fixtures import seeded_state_dict
from tests._golden import GOLDEN

pytestmark = pytest.mark.gpu

META = json.loads((GOLDEN / 'meta_val.json').read_text())
LOSS_RTOL = 1e-4
PROB_ATOL = 1e-3
LOG_ATOL = 2e-3
MASK_AGREE = 0.999


def _case(name):
    kw = dict(META['cases'][name])
    temperature = kw.pop('temperature')
    kw.pop('loss'), kw.pop('classes_present')
    feats, cls = V.synthetic_batch(**kw)
    g = np.load(GOLDEN / f'val_{name}.npz')
    return feats, cls, temperature, float(g['loss']), torch.from_numpy(g['pred']).long(), kw.get('half', False)


@pytest.mark.parametrize('name', sorted(META['cases']))
def test_cross_entropy_matches_reference_golden(name):
    from src.model.loss import CrossEntropy, propagate_clips
    feats, cls, temperature, want_loss, want_pred, half = _case(name)
    r, t, rc, tc = V.split_batch(feats, cls)
    gr, gt = (r.cuda().half(), t.cuda().half()) if half else (r.cuda(), t.cuda())   # f16-valued cases: one exact pass
    crit = CrossEntropy(temperature=temperature)
    onehot = torch.zeros(r.shape[0], r.shape[1], 22, 32, 32, device='cuda').scatter_(2, rc.cuda().unsqueeze(2), 1)
    loss, pred = crit(gr, gt, onehot, tc.cuda(), None, None, True)                   # the reference's calling convention
    loss_idx = crit(gr, gt, rc.cuda(), tc.cuda())                                     # class maps passed directly
    assert abs(float(loss) - float(loss_idx)) <= 1e-6 * float(loss)      # (nll_loss reduces with atomics)
    _, _, want_prob = V.cross_entropy(r, t, rc, tc, 22, temperature)
    prob = propagate_clips(gr, gt, rc.cuda(), 22, temperature).cpu()
    err = float((prob - want_prob).abs().max())
    agree = float((pred.cpu() == want_pred).float().mean())
    # the loss reads log(p + 1e-14): tiny probabilities must be right to relative, not absolute, accuracy
    dlog = float((torch.log(prob + 1e-14) - torch.log(want_prob + 1e-14)).abs().max())
    print(f'{name}: loss {float(loss):.7f} vs reference {want_loss:.7f}, max |dP| {err:.2e}, max |dlogP| {dlog:.2e}, '
          f'arg-max agreement {agree:.5f}, smallest true-class p {float(want_prob.gather(1, tc.reshape(tc.shape[0], 1, -1)).min()):.1e}')
    assert abs(float(loss) - want_loss) <= LOSS_RTOL * want_loss
    assert err <= PROB_ATOL and agree >= MASK_AGREE and dlog <= LOG_ATOL


@pytest.mark.parametrize('prec', ['f16', 'split3'])
def test_22_classes_with_spatial_prior(prec):
    """predict() proper (sampled references, Gaussian prior) with 22 classes: the wide-class instantiation of the
    index-label kernel against the oracle's predict."""
    from vosb200 import PREC_F16, PREC_SPLIT3, PropagationEngine, plan_refs
    T, d = 12, 22
    feats, _ = O.synthetic_sequence(T, 272, 400, 3, seed=77, feat_scale=0.30)
    if prec == 'f16':
        feats = feats.half().float()
    _, K, H_d, W_d = feats.shape
    P = H_d * W_d
    g = torch.Generator().manual_seed(3)
    coarse = torch.randint(0, d, (T, (H_d + 5) // 6, (W_d + 5) // 6), generator=g)
    cls = coarse.repeat_interleave(6, 1).repeat_interleave(6, 2)[:, :H_d, :W_d].reshape(T, P)
    hist = torch.stack([O.index_to_onehot(cls[f], d) for f in range(T)], 1)
    eng = PropagationEngine(max_pixels=P, ring_slots=48)
    eng.reset(H_d, W_d, 272, 400, d, PREC_F16 if prec == 'f16' else PREC_SPLIT3)
    gf = feats.cuda().half() if prec == 'f16' else feats.cuda()
    for f in range(T):
        eng.append(f, gf[f])
        eng.set_labels_index(f, cls[f].to(torch.uint8).cuda())
    for t in (1, 5, 11):
        refs, sig = plan_refs(t, 40, 9, 8.0, 21.0, False)
        out = eng.propagate(t, refs, sig, 1.0, False, write_labels=(t == 5))
        want = O.predict(feats[:t], feats[t], hist[:, :t], 8.0, 21.0, t, 40, 9, 1.0, False)
        got = out['prediction'].cpu()
        err = float((got - want).abs().max())
        agree = float((got.argmax(0) == want.argmax(0)).float().mean())
        print(f'd=22 {prec} t={t}: max |dP| {err:.2e}, arg-max agreement {agree:.5f}')
        assert err <= PROB_ATOL and agree >= MASK_AGREE
        assert np.array_equal(out['mask_lowres'].cpu().numpy(), got.argmax(0).numpy().astype(np.uint8))
        assert int(got.argmax(0).max()) > 14
        if t == 5:   # the written-back class bytes are what frame 11 reads as a reference: restore the teacher labels
            eng.set_labels_index(5, cls[5].to(torch.uint8).cuda())


def test_wide_class_error_paths():
    from vosb200 import PropagationEngine, VosPropError
    eng = PropagationEngine(max_pixels=1024, ring_slots=8)
    with pytest.raises(VosPropError):
        eng.reset(32, 32, 256, 256, 25)
    eng.reset(32, 32, 256, 256, 22)
    for f in range(2):
        eng.append(f, torch.zeros(256, 32, 32, device='cuda'))
    with pytest.raises(VosPropError):
        eng.set_labels_dense(0, torch.zeros(22, 1024, device='cuda'))
    eng.set_labels_index(0, torch.zeros(1024, dtype=torch.uint8))
    with pytest.raises(VosPropError):
        eng.propagate(1, [0], [0.0], topk=5)
    with pytest.raises(VosPropError):
        eng.propagate(1, [0], [0.0], probability_propagation=True)
    eng.propagate(1, [0], [0.0])
    eng.reset(16, 16, 128, 128, 22)          # narrower than the index-label kernel's 32-column groups
    for f in range(2):
        eng.append(f, torch.zeros(256, 16, 16, device='cuda'))
    eng.set_labels_index(0, torch.zeros(256, dtype=torch.uint8))
    with pytest.raises(VosPropError):
        eng.propagate(1, [0], [0.0])
    torch.cuda.synchronize()


class _Float64(torch.nn.Module):
    """Runs the stub embedder in fp64 on the GPU: the embeddings then equal the CPU fp32 ones of the golden run to
    fp32 rounding, independent of cuDNN's algorithm choice or TF32 (logits reach ~300 here, so 1e-3 relative noise in
    the embeddings would move the loss by more than the bar)."""

    def __init__(self, inner):
        super().__init__()
        self.inner = inner.double()

    def forward(self, x):
        return self.inner(x.double()).float()


def test_append_frames_equals_per_frame_appends():
    """vosprop_append_frames (a labelled clip in one call) leaves the engine in the state n x (append + set_labels_index)
    leaves it in: the propagated distribution is bit-identical."""
    from vosb200 import PropagationEngine, VosPropError
    feats, cls = V.synthetic_batch(1, T=6, seed=9, half=True)
    f, c = feats[0].cuda().half(), cls[0].cuda().to(torch.uint8)
    outs = []
    for batched in (False, True):
        eng = PropagationEngine(max_pixels=1024, ring_slots=12)
        eng.reset(32, 32, 256, 256, 22, 1)
        if batched:
            eng.append_frames(0, f[:5], c[:5])
        else:
            for t in range(5):
                eng.append(t, f[t])
                eng.set_labels_index(t, c[t])
        eng.append(5, f[5])
        outs.append(eng.propagate(5, list(range(5)), [0.0] * 5, write_labels=False)['prediction'].clone())
    assert torch.equal(outs[0], outs[1])
    with pytest.raises(VosPropError):
        eng.append_frames(0, torch.zeros(13, 256, 32, 32, device='cuda', dtype=torch.float16))     # more frames than ring slots
    with pytest.raises(ValueError):
        eng.append_frames(0, f[:2], c[:3].reshape(3, -1)[:, :100])
    torch.cuda.synchronize()


@pytest.mark.parametrize('layout', ['nchw', 'nhwc'])
@pytest.mark.parametrize('dtype', [torch.float16, torch.float32])
def test_batched_append_wraps_the_ring_like_single_appends(layout, dtype):
    """One launch for a batch of frames (grid z = frame) writes ring slot (first + i) % ring_slots of every frame, also
    across the wrap, from standard and channels-last sources: the propagated distribution is bit-identical to per-frame
    appends, and propagate_clip (which appends `engine.lookahead` frames ahead) to the per-frame loop."""
    from vosb200 import PropagationEngine
    from vosb200.engine import plan_refs
    torch.manual_seed(3)
    T, S = 15, 12
    f = (torch.randn(T, 256, 16, 40, device='cuda') * 0.3).to(dtype)
    if layout == 'nhwc':
        f = f.contiguous(memory_format=torch.channels_last)
    c = torch.randint(0, 3, (T, 16 * 40), device='cuda', dtype=torch.uint8)
    prec = 1 if dtype == torch.float16 else 0
    outs = []
    for batched in (False, True):
        eng = PropagationEngine(max_pixels=640, ring_slots=S)
        eng.reset(16, 40, 128, 320, 3, prec)
        if batched:
            eng.append_frames(0, f[:9], c[:9])
            eng.append_frames(9, f[9:T])                 # slots 9, 10, 11, 0, 1, 2: wraps
        else:
            for t in range(T):
                eng.append(t, f[t])
                if t < 9:
                    eng.set_labels_index(t, c[t])
        for t in range(9, T - 1):
            eng.set_labels_index(t, c[t])
        refs = [4, 5, 6, 8, 10, 11, 12, 13]              # frames 4 .. 13 are resident (3 was overwritten by 15 - 12)
        outs.append(eng.propagate(T - 1, refs, [8.0] * len(refs), write_labels=False)['prediction'].clone())
        eng.close()
    assert torch.equal(outs[0], outs[1])


def _loader(root, bs):
    from src.utils.datasets import TrainDataset
    ds = TrainDataset(root / 'JPEGImages/480p', root / 'Annotations/480p', frame_num=10, color_jitter=False)
    return torch.utils.data.DataLoader(ds, batch_size=bs, shuffle=False, num_workers=0, drop_last=True)


def test_step_matches_reference_step(tmp_path):
    """The mirror's step(mode='val') on the seeded synthetic dataset against the loss the reference's step() returned
    for it (same clips: tests/test_oracle_val.py checks the dataset bit for bit)."""
    from src.config import Config
    from src.model.loss import CrossEntropy
    from src.train import step
    sc = META['step']
    root = V.write_synthetic_dataset(tmp_path, sc['n_videos'], sc['n_frames'], sc['H'], sc['W'], sc['seed'])
    loader = _loader(root, sc['bs'])
    model = _Float64(V.StubEmbedder(sc['stub_seed'])).cuda()
    centroids = torch.Tensor(V.annotation_centroids()).float().cuda()
    old_dev = Config.DEVICE
    Config.DEVICE = torch.device('cuda')
    try:
        torch.manual_seed(sc['torch_seed'])
        loss = step(loader, model, CrossEntropy(temperature=1.0), None, 0, centroids, len(loader), mode='val')
    finally:
        Config.DEVICE = old_dev
    want = float(np.load(GOLDEN / 'val_step.npz')['loss'])
    print(f'step(): loss {loss:.7f} vs reference {want:.7f}')
    assert abs(loss - want) <= LOSS_RTOL * want


def test_validation_command_end_to_end(tmp_path):
    """`main.py validation` body on a dataset + checkpoint directory on disk; every checkpoint's loss must equal the
    oracle's loss on the embeddings the same network produces for the same clips."""
    from src.config import Config
    from src.model.vos_net import VOSNet
    from src.validation import validation_command_impl
    root = V.write_synthetic_dataset(tmp_path / 'data', 2, 11, 272, 336, seed=11)
    ckpts = tmp_path / 'ckpts'
    ckpts.mkdir()
    sds = {}
    for i, name in enumerate(('a.pth.tar', 'b.pth.tar')):
        sd = seeded_state_dict(VOSNet('resnet18', pretrained=False).state_dict())
        sd = {k: (v * (1.0 + 0.05 * i) if v.dtype.is_floating_point else v) for k, v in sd.items()}
        if i == 1:
            sd = {'module.' + k: v for k, v in sd.items()}      # a DataParallel checkpoint
        torch.save({'state_dict': sd}, ckpts / name)
        sds[name] = {k.replace('module.', ''): v for k, v in sd.items()}
    old = Config.DEVICE
    Config.DEVICE = torch.device('cuda')
    try:
        torch.manual_seed(99)
        got = validation_command_impl(str(root), str(ckpts), 2, 'cross_entropy', str(tmp_path / 'val.json'), 'resnet18', 0)
    finally:
        Config.DEVICE = old
    assert json.loads((tmp_path / 'val.json').read_text()) == got and sorted(got) == sorted(sds)
    # replay: same seed, same order of RNG draws (one loader pass per checkpoint, checkpoints in sorted order)
    centroids = torch.Tensor(V.annotation_centroids()).float()
    loader = _loader(root, 2)
    torch.manual_seed(99)
    for name in sorted(sds):
        net = VOSNet('resnet18', pretrained=False)
        net.load_state_dict(sds[name])
        net = net.cuda().eval()
        batches = [(img, ann) for img, ann, _ in loader]
        want, _ = V.validation_step(batches, lambda x: net(x.cuda()).float().cpu(), centroids)
        print(f'{name}: loss {got[name]:.6f} vs oracle on the same embeddings {want:.6f}')
        assert abs(got[name] - want) <= 1e-3 * want
