"""GPU parity: the CUDA path (through the C ABI) against the reference goldens and the CPU oracle.

Bars (BASELINE.json north_star): masks bit-exact except documented near-tie pixels, >= 99.9 %
agreement; probabilities within 1e-3 absolute of the fp32 reference (bf16x3 tensor-core kernel)."""
import numpy as np
import pytest
import torch

from oracle import propagation_oracle as O
from tests import _golden as G

pytestmark = pytest.mark.gpu

PROB_ATOL = 1e-3
MASK_AGREE = 0.999


def _engine(max_pixels, ring_slots=48):
    from vosb200 import PropagationEngine
    return PropagationEngine(max_pixels=max_pixels, ring_slots=ring_slots)


def _kernels():
    # 'tc' = product dispatch (index-label kernel when labels are class ids and W_d >= 32, else the
    # general kernel); 'tc_dense' forces the general tensor-core kernel; 'simt' = fp32 checker
    from vosb200 import KERNEL_SIMT, KERNEL_TC, KERNEL_TC_DENSE
    return {'tc': KERNEL_TC, 'tc_dense': KERNEL_TC_DENSE, 'simt': KERNEL_SIMT}


KERNELS = ['simt', 'tc_dense', 'tc']


@pytest.mark.parametrize('kernel', KERNELS)
@pytest.mark.parametrize('name', G.SEQ_NAMES)
def test_golden_sequences(name, kernel):
    """Whole clips against outputs of the reference's real inference_single."""
    from vosb200.sequence import propagate_clip
    feats, first, run = G.sequence_inputs(name)
    masks_ref, preds_ref = G.sequence_golden(name)
    eng = _engine(feats.shape[2] * feats.shape[3])
    masks, preds = propagate_clip(eng, feats.cuda(), first, kernel=_kernels()[kernel], return_predictions=True, **run)
    masks, preds = masks.cpu().numpy(), preds.cpu().numpy()
    agree = float((masks == masks_ref).mean())
    err = float(np.abs(preds - preds_ref).max())
    print(f'{name}/{kernel}: mask agreement {agree:.6f}, max |dP| {err:.3e}')
    assert agree >= MASK_AGREE
    if agree == 1.0 or run['probability_propagation']:
        # label drift after a near-tie flip changes later frames' inputs; only compare probabilities
        # when both sides propagated the same labels
        assert err <= PROB_ATOL


@pytest.mark.parametrize('kernel', KERNELS)
@pytest.mark.parametrize('tag', G.SEQ16_NAMES)
def test_golden_sequences_16bit_single_pass(tag, kernel):
    """fp16 / bf16 embeddings (what VOSNet emits under autocast): ONE tensor-core pass, exact products.
    Goldens: the reference's inference_single run in fp32 on the same 16-bit-valued embeddings."""
    from vosb200 import PREC_BF16, PREC_F16
    from vosb200.sequence import propagate_clip
    feats, first, run = G.sequence16_inputs(tag)
    masks_ref, preds_ref = G.sequence16_golden(tag)
    eng = _engine(feats.shape[2] * feats.shape[3])
    masks, preds = propagate_clip(eng, feats.cuda(), first, kernel=_kernels()[kernel], return_predictions=True, **run)
    assert eng.precision == (PREC_F16 if feats.dtype == torch.float16 else PREC_BF16)
    masks, preds = masks.cpu().numpy(), preds.cpu().numpy()
    agree = float((masks == masks_ref).mean())
    err = float(np.abs(preds - preds_ref).max())
    print(f'{tag}/{kernel}: mask agreement {agree:.6f}, max |dP| {err:.3e}')
    assert agree >= MASK_AGREE
    if agree == 1.0 or run['probability_propagation']:
        assert err <= PROB_ATOL


@pytest.mark.parametrize('kernel', ['simt', 'tc_dense'])
def test_golden_predict_cases_teacher_forced(kernel):
    """Stand-alone predict() calls with externally supplied label histories (teacher forcing:
    no drift), incl. duplicated references, int32 first-frame labels, both sigma branches."""
    from vosb200 import plan_refs
    feats, hist, prob_hist = G.predict_case_inputs()
    z = np.load(G.GOLDEN / 'predict_cases.npz')
    T, K, H_d, W_d = feats.shape
    P = H_d * W_d
    eng = _engine(P, ring_slots=64)
    gf = feats.cuda()
    for key in z.files:
        c = G.parse_case(key)
        lab = prob_hist if c['prob'] else hist
        t = c['t']
        refs, sig = plan_refs(t, c['frame_range'], c['ref_num'], 8.0, 21.0, c['prob'])
        eng.reset(H_d, W_d, H_d * 8, W_d * 8, lab.shape[0])
        for f in sorted(set(refs)):
            eng.append(f, gf[f])
            eng.set_labels_dense(f, lab[:, f].cuda())
        eng.append(t, gf[t])
        out = eng.propagate(t, refs, sig, c['temperature'], c['prob'], write_labels=False,
                            kernel=_kernels()[kernel], want_fullres=False)
        got = out['prediction'].cpu().numpy()
        err = float(np.abs(got - z[key]).max())
        print(f'{key}/{kernel}: max |dP| {err:.3e}')
        assert err <= PROB_ATOL, key
        near_tie = np.sort(z[key], axis=0)[-1] - np.sort(z[key], axis=0)[-2] < 2 * PROB_ATOL
        same = got.argmax(0) == z[key].argmax(0)
        assert (same | near_tie).all(), key
        assert np.array_equal(out['mask_lowres'].cpu().numpy(), got.argmax(0).astype(np.uint8))


@pytest.mark.parametrize('prec', ['split3', 'f16'])
@pytest.mark.parametrize('kernel', KERNELS)
def test_480p_against_oracle(kernel, prec):
    """480p (60x107 = 6420 pixels, 51 tiles, ragged last tile) vs the CPU oracle, teacher forced.
    Labels are class ids ('tc' -> index-label kernel) with a random history: every 32-pixel chunk is
    class-mixed, the hardest case for the ballot path."""
    from vosb200 import plan_refs
    T = 18
    from vosb200 import PREC_F16, PREC_SPLIT3
    feats, first = O.synthetic_sequence(T, 480, 854, 2, seed=31, feat_scale=0.30)
    if prec == 'f16':
        feats = feats.half().float()     # the oracle sees exactly the values the engine stores
    _, K, H_d, W_d = feats.shape
    P = H_d * W_d
    low, d = O.first_frame_labels(first)
    g = torch.Generator().manual_seed(7)
    cls = torch.randint(0, d, (T, P), generator=g)
    cls[1::2] = torch.where(torch.rand(T, P, generator=g)[1::2] < 0.97, cls[1::2] * 0 + 1, cls[1::2])  # mostly homogeneous frames
    cls[0] = low
    hist = torch.stack([O.index_to_onehot(cls[f], d) for f in range(T)], 1)
    eng = _engine(P)
    eng.reset(H_d, W_d, 480, 854, d, PREC_F16 if prec == 'f16' else PREC_SPLIT3)
    gf = feats.cuda().half() if prec == 'f16' else feats.cuda()
    for f in range(T):
        eng.append(f, gf[f])
        eng.set_labels_index(f, cls[f].to(torch.uint8).cuda())
    for t in (1, 9, 17):
        refs, sig = plan_refs(t, 40, 9, 8.0, 21.0, False)
        out = eng.propagate(t, refs, sig, 1.0, False, write_labels=False, kernel=_kernels()[kernel])
        want = O.predict(feats[:t], feats[t], hist[:, :t], 8.0, 21.0, t, 40, 9, 1.0, False, chunk=1024)
        got = out['prediction'].cpu()
        err = float((got - want).abs().max())
        agree = float((got.argmax(0) == want.argmax(0)).float().mean())
        print(f'480p t={t}/{kernel}/{prec}: max |dP| {err:.3e}, argmax agreement {agree:.6f}')
        assert err <= PROB_ATOL
        assert agree >= MASK_AGREE
        full = O.upsample_mask(out['mask_lowres'].cpu().long(), H_d, W_d, 480, 854)
        assert torch.equal(out['mask'].cpu().long(), full)


def test_tc_matches_simt_bitwise_masks_on_clip():
    """The tensor-core kernel and the fp32 CUDA-core checker share everything but the logits."""
    from vosb200.sequence import propagate_clip
    feats, first = O.synthetic_sequence(12, 240, 432, 3, seed=41, feat_scale=0.30)
    eng = _engine(feats.shape[2] * feats.shape[3])
    k = _kernels()
    m_tc, p_tc = propagate_clip(eng, feats.cuda(), first, kernel=k['tc'], return_predictions=True)
    m_si, p_si = propagate_clip(eng, feats.cuda(), first, kernel=k['simt'], return_predictions=True)
    agree = float((m_tc == m_si).float().mean())
    err = float((p_tc - p_si).abs().max())
    print(f'tc vs simt: mask agreement {agree:.6f}, max |dP| {err:.3e}')
    assert agree >= MASK_AGREE


def test_error_paths():
    from vosb200 import VosPropError
    eng = _engine(240, ring_slots=8)
    with pytest.raises(VosPropError):
        eng.geom = (12, 20, 96, 160, 3)
        eng.append(0, torch.zeros(256, 12, 20, device='cuda'))  # reset() never called on the C side
    eng.reset(12, 20, 96, 160, 3)
    eng.append(0, torch.zeros(256, 12, 20, device='cuda'))
    eng.append(1, torch.zeros(256, 12, 20, device='cuda'))
    with pytest.raises(VosPropError):   # labels of frame 0 never set
        eng.propagate(1, [0], [8.0])
    eng.set_labels_index(0, torch.zeros(240, dtype=torch.uint8))
    eng.propagate(1, [0], [8.0])
    with pytest.raises(VosPropError):   # frame 5 not resident
        eng.propagate(1, [5], [8.0])
    with pytest.raises(VosPropError):   # negative temperature unsupported (documented)
        eng.propagate(1, [0], [8.0], temperature=-1.0)
    from vosb200 import PREC_F16
    eng.reset(12, 20, 96, 160, 3, PREC_F16)
    with pytest.raises(VosPropError):   # fp32 embeddings into an fp16 memory would be rounded: refused
        eng.append(0, torch.zeros(256, 12, 20, device='cuda'))
    eng.append(0, torch.zeros(256, 12, 20, device='cuda', dtype=torch.float16))
    with pytest.raises(VosPropError):
        eng.reset(12, 20, 96, 160, 25)  # too many classes (15..24 run on the index-label kernel only)
    with pytest.raises(VosPropError):
        eng.reset(120, 200, 960, 1600, 3)  # beyond capacity
    torch.cuda.synchronize()


def test_fused_backbone_matches_module_graph():
    """BN-folded cuDNN-fused trunk vs the stock VOSNet module graph (fp32) on the same weights."""
    from oracle.fixtures import seeded_state_dict
    from src.model.vos_net import VOSNet
    from vosb200.fused_backbone import FusedVOSNet
    net = VOSNet('resnet50', pretrained=False).eval()
    net.load_state_dict(seeded_state_dict(net.state_dict()))
    net = net.cuda()
    x = torch.randn(2, 3, 96, 160, device='cuda', generator=torch.Generator(device='cuda').manual_seed(0))
    with torch.no_grad():
        want = net(x)
        got = FusedVOSNet(net)(x).float()
    rel = float((got - want).abs().max() / want.abs().max())
    print(f'fused backbone: max rel err {rel:.3e}')
    assert got.shape == want.shape and rel < 2e-2      # fp16 convolutions


def test_clip_segmenter_end_to_end_matches_stepwise_engine():
    """The public e2e call (pinned host frames -> masks on host) equals feeding its own embeddings through
    the engine frame by frame."""
    from oracle.fixtures import seeded_state_dict
    from src.model.vos_net import VOSNet
    from vosb200 import synthetic
    from vosb200.pipeline import ClipSegmenter
    from vosb200.sequence import propagate_clip
    net = VOSNet('resnet50', pretrained=False).eval()
    net.load_state_dict(seeded_state_dict(net.state_dict()))
    seg = ClipSegmenter(net, backbone_batch=4)
    frames, first = synthetic.clip_frames(9, 256, 320, 2, seed=3, device='cuda')
    masks = seg.segment(frames, first)
    assert masks.shape == (8, 256, 320) and masks.dtype == torch.uint8 and not masks.is_cuda
    feats = torch.cat([seg.embed(frames[i:i + 4].cuda()) for i in range(0, 9, 4)]).float()
    want = propagate_clip(_engine(feats.shape[2] * feats.shape[3]), feats, first.cuda())
    agree = float((masks.cuda() == want).float().mean())
    print(f'ClipSegmenter vs stepwise: {agree:.6f}')
    assert agree >= MASK_AGREE


def test_clip_segmenter_async_copy_back_equals_synchronous_call():
    """segment(sync=False) returns while the masks are still on their way to the host (own D2H stream, behind the next
    clip's compute); wait() / a device synchronise completes them.  Several clips queued back to back -- uint8 frames and
    fp32 normalised ones -- give the masks of one synchronous call per clip, bit for bit."""
    from oracle.fixtures import seeded_state_dict
    from src.model.vos_net import VOSNet
    from vosb200 import synthetic
    from vosb200.pipeline import ClipSegmenter
    net = VOSNet('resnet50', pretrained=False).eval()
    net.load_state_dict(seeded_state_dict(net.state_dict()))
    seg = ClipSegmenter(net, backbone_batch=4)
    clips = [synthetic.clip_frames(7 + i, 256, 320, 1 + i % 3, seed=11 + i, device='cuda', raw=(i % 2 == 0)) for i in range(4)]
    want = [seg.segment(f, first).clone() for f, first in clips]
    outs = [torch.empty((f.shape[0] - 1, 256, 320), dtype=torch.uint8, pin_memory=True) for f, _ in clips]
    for (f, first), out in zip(clips, outs):
        got = seg.segment(f, first, out=out, sync=False)
        assert got is out
    seg.wait()                       # the last clip's copy: D2H copies run in order on one stream, so all of them
    for i, (w, o) in enumerate(zip(want, outs)):
        assert torch.equal(w, o), f'clip {i}'
    # a device-side first annotation takes the class count from the device (one host sync), same masks
    f, first = clips[1]
    assert torch.equal(seg.segment(f, first.cuda()), want[1])


def test_clip_segmenter_pool_equals_one_clip_at_a_time():
    """ClipSegmenterPool: clips in flight on several lanes (own engine + stream each) give the masks of one synchronous
    ClipSegmenter.segment call per clip, bit for bit -- for 1, 2 and 3 lanes, more clips than lanes, mixed sizes."""
    from oracle.fixtures import seeded_state_dict
    from src.model.vos_net import VOSNet
    from vosb200 import synthetic
    from vosb200.pipeline import ClipSegmenter, ClipSegmenterPool
    net = VOSNet('resnet50', pretrained=False).eval()
    net.load_state_dict(seeded_state_dict(net.state_dict()))
    sizes = [(256, 320), (192, 256), (256, 320), (200, 264), (256, 320)]
    clips = [synthetic.clip_frames(6 + 2 * i, h, w, 1 + i % 3, seed=21 + i, device='cuda', raw=(i % 2 == 1))
             for i, (h, w) in enumerate(sizes)]
    one = ClipSegmenter(net, backbone_batch=4)
    want = [one.segment(f, first).clone() for f, first in clips]
    for lanes in (1, 2, 3):
        pool = ClipSegmenterPool(net, lanes=lanes, backbone_batch=4)
        got = pool.segment_many(clips)
        assert len(got) == len(clips)
        for i, (w, g) in enumerate(zip(want, got)):
            assert g.shape == w.shape and torch.equal(w, g), f'lanes {lanes}, clip {i}'
        outs = [torch.empty_like(w).pin_memory() for w in want]
        pool.segment_many(clips, outs=outs, sync=False)
        pool.wait()
        assert all(torch.equal(w, o) for w, o in zip(want, outs)), f'lanes {lanes}, preallocated outputs'
        pool.close()
