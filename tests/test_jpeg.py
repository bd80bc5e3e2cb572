"""JPEG front end (include/vos_jpeg.h): the reference decodes every frame with Pillow
(`Image.open(BytesIO(bytes)).convert('RGB')`, /root/reference/src/utils/datasets.py:141-143).  Pillow is that call's
implementation and is present wherever these tests run, so it is the golden source itself: every case encodes a seeded image
with Pillow, decodes it with Pillow, and demands the same bytes from
  * the numpy restatement (oracle/jpeg_oracle.py)                       -- CPU
  * the library's host path (Huffman stage + vosjpeg_reconstruct_host)  -- CPU, through the C ABI
  * the library's device path (vosjpeg_reconstruct)                     -- GPU, through the C ABI
Bar: bit-exact (byte work)."""
import ctypes as C
import io
import re
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
from PIL import Image

REPO = Path(__file__).resolve().parent.parent
for p in (str(REPO), str(REPO / 'semi-supervised-vos_b200')):
    if p not in sys.path:
        sys.path.insert(0, p)

from oracle import jpeg_oracle as O  # noqa: E402
from vosb200 import _capi as capi  # noqa: E402
from vosb200 import jpeg as J  # noqa: E402


def picture(h, w, seed=0, noise=20.0):
    rs = np.random.RandomState(seed)
    y, x = np.mgrid[0:h, 0:w]
    base = np.stack([128 + 100 * np.sin(x / 7.0 + y / 13.0), 128 + 90 * np.cos(x / 5.0 - y / 9.0), 255.0 * ((x // 9 + y // 11) % 2)], -1)
    return np.clip(base + rs.randn(h, w, 3) * noise, 0, 255).astype(np.uint8)


def encode(a, **kw):
    b = io.BytesIO()
    Image.fromarray(a).save(b, format='JPEG', **kw)
    return b.getvalue()


def pillow(data):
    return np.asarray(Image.open(io.BytesIO(data)).convert('RGB'))


SIZES = [(64, 80), (37, 53), (16, 16), (100, 75), (8, 24), (33, 130), (3, 5), (1, 1), (17, 2), (2, 17)]
CASES = [(hw, q, sub) for hw in SIZES for q in (30, 90, 100) for sub in (0, 1, 2)]
EXTRA = [dict(quality=85, optimize=True), dict(quality=85, restart_marker_blocks=3), dict(quality=85, restart_marker_rows=1),
         dict(quality=60, subsampling=2, restart_marker_blocks=1), dict(quality=95, qtables='web_high')]


def all_streams():
    for k, ((h, w), q, sub) in enumerate(CASES):
        yield f'{h}x{w} q{q} sub{sub}', encode(picture(h, w, seed=k), quality=q, subsampling=sub)
    for k, kw in enumerate(EXTRA):
        yield str(kw), encode(picture(40, 56, seed=100 + k), **kw)
    yield 'grey', encode(picture(40, 56, seed=7)[..., 0], quality=85)


def test_oracle_reproduces_pillow():
    for name, data in all_streams():
        assert np.array_equal(O.decode(data), pillow(data)), name


def test_header_declares_what_the_library_exports():
    header = (REPO / 'include' / 'vos_jpeg.h').read_text()
    declared = set(re.findall(r'\b(vosjpeg_[a-z_]+)\s*\(', header))
    assert declared == set(J.EXPORTS), declared ^ set(J.EXPORTS)
    lib = C.CDLL(str(capi.LIB_PATH))
    for name in declared:
        assert hasattr(lib, name), name
    assert C.sizeof(J.Info) == 520          # vosjpeg_info: the layout both sides were built with


def test_host_path_reproduces_pillow_and_the_oracle_coefficients():
    for name, data in all_streams():
        info = J.parse(data)
        coef = J.entropy_decode(data, info, pinned=False)
        h = O.parse(data)
        O.entropy_decode(data, h)
        assert (info.width, info.height, info.n_comp) == (h.width, h.height, len(h.comps)), name
        for c, comp in enumerate(h.comps):
            got = coef[info.coef_offset[c]: info.coef_offset[c] + comp.coef.size].numpy().reshape(comp.coef.shape)
            assert np.array_equal(got, comp.coef), f'{name}: coefficients of component {c}'
            assert np.array_equal(np.array(info.quant[c][:]), h.qt[comp.tq]), name
        assert np.array_equal(J.reconstruct(info, coef).numpy(), pillow(data)), name


def test_480p_frame_on_the_host():
    data = encode(picture(480, 854, seed=3, noise=6.0), quality=90)
    assert np.array_equal(J.decode(data).numpy(), pillow(data))


def test_flavours_outside_the_path_are_refused_not_decoded():
    a = picture(40, 56)
    for data in (encode(a, quality=85, progressive=True), io.BytesIO()):
        if isinstance(data, io.BytesIO):
            Image.fromarray(a).convert('CMYK').save(data, format='JPEG')
            data = data.getvalue()
        with pytest.raises(J.Unsupported):
            J.parse(data)
    swapped = bytearray(encode(a, quality=85, subsampling=2))       # luma 1x1 under 2x2 chroma: no encoder writes it, the device
    sof = swapped.index(b'\xff\xc0')                                 # stage reads luma at full resolution -> refused
    swapped[sof + 11], swapped[sof + 14], swapped[sof + 17] = 0x11, 0x22, 0x22
    with pytest.raises(J.Unsupported):
        J.parse(bytes(swapped))
    with pytest.raises(J.JpegError):
        J.parse(b'not a jpeg at all')
    with pytest.raises(J.JpegError):
        J.parse(encode(a)[:100])


def _without_jfif(data, ids=None):
    """The stream without its APP0 JFIF segment, optionally with the component identifiers of SOF0 / SOS replaced."""
    assert data[2:4] == b'\xff\xe0'
    out = bytearray(data[:2] + data[4 + ((data[4] << 8) | data[5]):])
    if ids is not None:
        sof = out.index(b'\xff\xc0')
        sos = out.index(b'\xff\xda')
        for c, cid in enumerate(ids):
            out[sof + 10 + 3 * c] = cid
            out[sos + 5 + 2 * c] = cid
    return bytes(out)


def test_colour_space_guess_follows_libjpeg():
    """jdapimin.c: no JFIF / Adobe marker -> components 1, 2, 3 (or anything unknown) are YCbCr, but 'R', 'G', 'B' are RGB data
    that Pillow returns unconverted: that file is refused (the loader keeps Pillow for it), the others decode to Pillow's bytes."""
    data = encode(picture(40, 56, seed=2), quality=90, subsampling=0)
    plain = _without_jfif(data)
    assert np.array_equal(J.decode(plain).numpy(), pillow(plain)) and np.array_equal(O.decode(plain), pillow(plain))
    odd = _without_jfif(data, ids=(7, 8, 9))
    assert np.array_equal(J.decode(odd).numpy(), pillow(odd))
    rgb = _without_jfif(data, ids=(ord('R'), ord('G'), ord('B')))
    assert not np.array_equal(pillow(rgb), pillow(plain))           # Pillow really treats it differently
    with pytest.raises(J.Unsupported):
        J.parse(rgb)
    with pytest.raises(O.Unsupported):
        O.parse(rgb)


def test_damaged_streams_never_crash():
    """Truncations and flipped bytes: an error or some picture, never a fault (bounds are the library's business)."""
    rs = np.random.RandomState(5)
    good = encode(picture(48, 64, seed=9), quality=80, restart_marker_blocks=4)
    for trial in range(300):
        data = bytearray(good)
        if trial % 3 == 0:
            data = data[:rs.randint(2, len(data))]
        else:
            for _ in range(rs.randint(1, 6)):
                data[rs.randint(2, len(data))] = rs.randint(256)
        data = bytes(data)
        try:
            info = J.parse(data)
            if info.coef_count > (1 << 24):         # a flipped size field: nothing to learn from decoding a huge blank
                continue
            coef = J.entropy_decode(data, info, pinned=False)
            J.reconstruct(info, coef)
        except J.JpegError:
            pass


def test_loader_items_round_trip_on_the_host():
    datas = [encode(picture(72, 104, seed=k), quality=88) for k in range(3)]
    items = torch.stack([J.pack_item(d) for d in datas])
    frames = J.unpack_items(items, 'cpu')
    for f, d in zip(frames, datas):
        assert np.array_equal(f.numpy(), pillow(d))


def test_dataset_ships_coefficients_and_falls_back_to_pillow(tmp_path):
    from src.utils.datasets import InferenceDataset
    root = tmp_path / 'JPEGImages'
    (root / 'vid').mkdir(parents=True)
    a = picture(64, 96, seed=4)
    (root / 'vid' / '00000.jpg').write_bytes(encode(a, quality=90))
    (root / 'vid' / '00001.jpg').write_bytes(encode(a, quality=90, progressive=True))
    ds = InferenceDataset(str(root), disable=True, raw='coef')
    item0, video = ds[0]
    assert video == 'vid' and item0.dtype == torch.int16 and item0.dim() == 1
    assert np.array_equal(J.unpack_items(item0[None], 'cpu')[0].numpy(), pillow((root / 'vid' / '00000.jpg').read_bytes()))
    item1, _ = ds[1]                                   # progressive: Pillow's frame, as with raw=True
    assert item1.dtype == torch.uint8 and np.array_equal(item1.numpy(), pillow((root / 'vid' / '00001.jpg').read_bytes()))
    flip = InferenceDataset(str(root), disable=True, raw='coef', inference_strategy='hor-flip')
    (x, x_flipped), _ = flip[0]                        # strategies that transform the decoded image keep Pillow
    assert x.dtype == torch.uint8 and x_flipped.dtype == torch.uint8


@pytest.mark.gpu
def test_device_path_reproduces_pillow():
    for name, data in all_streams():
        info = J.parse(data)
        coef = J.entropy_decode(data, info)
        rgb = J.reconstruct(info, coef.cuda())
        assert rgb.is_cuda and np.array_equal(rgb.cpu().numpy(), pillow(data)), name


@pytest.mark.gpu
def test_480p_frames_on_the_device_and_through_the_loop_input_stage():
    from src.utils.inference_utils import _to_device
    from vosb200 import normalize_frames
    datas = [encode(picture(480, 854, seed=k, noise=4.0 + 6 * k), quality=q, subsampling=s) for k, (q, s) in enumerate(((90, 2), (75, 1), (95, 0)))]
    for d in datas:
        assert np.array_equal(J.decode(d, device='cuda').cpu().numpy(), pillow(d))
    # one batch = one launch pair; the frames differ in quality, i.e. in their quantisation tables (read from each item's header)
    same = [encode(picture(480, 854, seed=10 + k, noise=5.0), quality=q) for k, q in enumerate((90, 60, 75, 95, 90))]
    items = torch.stack([J.pack_item(d) for d in same])
    got = _to_device(items)                                             # what the loops feed the network
    want = normalize_frames(torch.from_numpy(np.stack([pillow(d) for d in same])).cuda(), torch.float32)
    assert got.shape == want.shape and torch.equal(got, want)



def test_batch_of_files_on_the_librarys_threads():
    """vosjpeg_decode_files_host: the host stage for a batch of files on several threads == one pack_item per file; a flavour
    outside the path and a frame larger than planned for are reported per file, the others are unaffected."""
    datas = [encode(picture(72, 104, seed=k), quality=80 + k) for k in range(9)]
    datas[3] = encode(picture(72, 104, seed=3), quality=85, progressive=True)
    datas[6] = encode(picture(144, 208, seed=6), quality=85)
    capacity = J.item_length(datas[0])
    for threads in (1, 3, 16):
        buf, status = J.pack_items_threaded(datas, capacity, threads)
        assert status == [0, 0, 0, J.ERR_UNSUPPORTED, 0, 0, J.ERR_UNSUPPORTED, 0, 0]
        for i, st in enumerate(status):
            if st == 0:
                assert J.item_values(buf[i]) == capacity and torch.equal(buf[i], J.pack_item(datas[i])), (threads, i)


def test_worker_processes_ship_coefficients(tmp_path):
    """The command's loader: DataLoader worker processes run the host stage (the library is loaded in the workers, no CUDA call)
    and the items that come back decode to Pillow's frames, in order."""
    from src.utils.datasets import InferenceDataset
    root = tmp_path / 'JPEGImages'
    (root / 'v').mkdir(parents=True)
    for t in range(10):
        (root / 'v' / f'{t:05d}.jpg').write_bytes(encode(picture(96, 128, seed=t), quality=90, subsampling=t % 3))
    ds = InferenceDataset(str(root), disable=True, raw='coef')
    loader = torch.utils.data.DataLoader(ds, batch_size=1, shuffle=False, num_workers=2, prefetch_factor=2)
    seen = 0
    for i, (item, (video,)) in enumerate(loader):
        assert video == 'v' and item.dtype == torch.int16
        assert np.array_equal(J.unpack_items(item, 'cpu')[0].numpy(), pillow((root / 'v' / f'{i:05d}.jpg').read_bytes())), i
        seen += 1
    assert seen == 10
