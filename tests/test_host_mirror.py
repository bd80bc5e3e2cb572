"""CPU checks of the host-side mirror of the reference interface (semi-supervised-vos_b200/src)."""
import inspect
import json

import numpy as np
import pytest
import torch

from oracle import propagation_oracle as O
from oracle.fixtures import seeded_state_dict
from tests import _golden as G


def test_vosnet_matches_reference_bit_exact():
    from src.model.vos_net import VOSNet
    meta = G.META['vosnet_forward']
    net = VOSNet('resnet50', pretrained=False).eval()
    sd = net.state_dict()
    assert {k: list(v.shape) for k, v in sd.items()} == meta['keys']      # the reference's checkpoints load
    net.load_state_dict(seeded_state_dict(sd))
    x = torch.randn(*meta['input_shape'], generator=torch.Generator().manual_seed(meta['input_seed']))
    with torch.no_grad():
        y = net(x)
    assert np.array_equal(y.numpy(), np.load(G.GOLDEN / 'vosnet_forward.npz')['y'])


def test_vosnet_variants_and_checkpoint_loader(tmp_path):
    from src.model.vos_net import VOSNet
    from src.utils.utils import load_model
    n18 = VOSNet('resnet18', pretrained=False).eval()
    with torch.no_grad():
        assert n18(torch.zeros(1, 3, 64, 96)).shape == (1, 256, 8, 12)
    net = VOSNet('resnet50', pretrained=False)
    ck = tmp_path / 'a.pth.tar'
    torch.save({'epoch': 3, 'state_dict': net.state_dict()}, ck)                       # train.py:144-151 format
    load_model(VOSNet('resnet50', pretrained=False), str(ck))
    torch.save({'module.' + k: v for k, v in net.state_dict().items()}, ck)            # DataParallel-prefixed, bare
    load_model(VOSNet('resnet50', pretrained=False), str(ck))
    with pytest.raises(NotImplementedError):
        VOSNet('vgg')


def test_first_frame_lowres_matches_reference_get_labels():
    from vosb200.sequence import first_frame_lowres, lowres_dims
    z = np.load(G.GOLDEN / 'first_frame_labels.npz')
    for key in z.files:
        if not key.startswith('first_'):
            continue
        first = torch.from_numpy(z[key].astype(np.int64))
        H_d, W_d = lowres_dims(*first.shape)
        assert (H_d, W_d) == O.lowres_dims(*first.shape)
        assert np.array_equal(first_frame_lowres(first, H_d, W_d).numpy(), z['low_' + key[6:]])


def test_spatial_prior_descriptor_materialises_to_the_reference_matrix():
    from src.model.predict import SpatialPrior, _sigma_of, get_spatial_weight
    z = np.load(G.GOLDEN / 'spatial_weight.npz')
    for key in z.files:
        _, h, w, s = key.split('_')
        prior = get_spatial_weight((int(h), int(w)), float(s))
        assert isinstance(prior, SpatialPrior)
        assert np.array_equal(prior.materialize().numpy(), z[key])
        # an explicit (P,P) matrix handed to predict() is accepted: sigma is recovered from it
        assert abs(_sigma_of(torch.from_numpy(z[key]), int(w)) - float(s)) < 1e-3 * float(s)


def test_png_writer_round_trip(tmp_path):
    from PIL import Image
    from src.utils.utils import save_predictions
    pal = [0, 0, 0, 128, 0, 0, 0, 128, 0] + [0] * (768 - 9)
    masks = np.random.default_rng(0).integers(0, 3, size=(3, 24, 40)).astype(np.uint8)
    save_predictions(masks, pal, str(tmp_path), 'vid')
    for i in range(3):
        img = Image.open(tmp_path / 'vid' / f'{i + 1:05d}.png')
        assert img.mode == 'P' and img.getpalette()[:9] == pal[:9]
        assert np.array_equal(np.asarray(img), masks[i])


def test_mirror_signatures_match_the_reference():
    """Parameter names of the drop-in callables (recorded from the reference by make_golden.py)."""
    import src.inference as inf
    import src.model.predict as pr
    import src.utils.inference_utils as iu
    ours = {'predict': pr.predict, 'sample_frames': pr.sample_frames, 'prepare_first_frame': pr.prepare_first_frame,
            'get_labels': pr.get_labels, 'get_spatial_weight': pr.get_spatial_weight,
            'inference_single': iu.inference_single, 'inference_command_impl': inf.inference_command_impl}
    want = G.META['signatures']
    for name, fn in ours.items():
        assert list(inspect.signature(fn).parameters) == want[name], name
    assert sorted(p.name for p in inf.inference_command.params) == sorted(want['inference_command_options'])


def test_cli_surface():
    from click.testing import CliRunner
    import main
    r = CliRunner().invoke(main.cli, ['--help'])
    assert 'inference' in r.output and 'validation' in r.output
    r = CliRunner().invoke(main.cli, ['validation', '--help'])
    for opt in ('--data', '--checkpoints', '--bs', '--loss', '--miner', '--margin', '--loss_weight', '--output'):
        assert opt in r.output                       # the reference's options (src/validation.py:30-42)
    r = CliRunner().invoke(main.cli, ['validation', '-d', '.', '-c', '.', '-o', 'x.json', '--loss', 'triplet'])
    assert r.exit_code != 0 and 'not built' in r.output


def test_lpt_assignment_is_balanced_and_deterministic():
    from vosb200.shard import assign_lpt, imbalance, sequence_cost
    rng = np.random.default_rng(3)
    lens = rng.integers(34, 105, size=30)              # DAVIS-2017-val-shaped lengths
    costs = [sequence_cost(int(n), 6420) for n in lens]
    a = assign_lpt(costs, 8)
    assert sorted(i for r in a for i in r) == list(range(30))
    assert a == assign_lpt(costs, 8)
    assert imbalance(costs, a) < 1.10
    assert assign_lpt(costs, 1) == [list(range(30))]
    assert sequence_cost(10, 100, 9) == (1 + 2 + 3 + 4 + 5 + 6 + 7 + 8 + 9) * 100.0 ** 2
