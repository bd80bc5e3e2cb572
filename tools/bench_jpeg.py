"""JPEG front end (include/vos_jpeg.h) at 480p: host Huffman stage next to Pillow's full decode (one core each), device stage
(de-quantise + inverse DCT + up-sample + colour, two kernels) timed with CUDA events against the HBM roofline.
Algorithmic bytes of the device stage per frame: coefficients read (2 B each) + sample planes written and read once + RGB written."""
import io
import json
import sys
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
for p in (str(REPO), str(REPO / 'semi-supervised-vos_b200')):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from PIL import Image  # noqa: E402

from vosb200 import jpeg as J  # noqa: E402


def main():
    peaks = REPO / 'MEASURED_PEAKS.json'
    hbm = None
    if peaks.is_file():
        pk = json.loads(peaks.read_text())
        hbm = pk.get('hbm_gbs_burst') or pk.get('hbm_gbs') or pk.get('hbm_copy_gbs')
    rs = np.random.RandomState(0)
    H, W = 480, 854
    base = np.asarray(Image.fromarray(rs.randint(0, 256, (H // 32 + 1, W // 32 + 1, 3)).astype(np.uint8)).resize((W, H), Image.BILINEAR))
    for quality, sub in ((90, 2), (90, 0)):
        b = io.BytesIO()
        Image.fromarray(base).save(b, format='JPEG', quality=quality, subsampling=sub)
        data = b.getvalue()
        info = J.parse(data)
        coef = J.entropy_decode(data, info)
        t0 = time.perf_counter()
        for _ in range(30):
            J.entropy_decode(data, info, out=coef)
        t_huff = (time.perf_counter() - t0) / 30
        t0 = time.perf_counter()
        for _ in range(30):
            np.asarray(Image.open(io.BytesIO(data)).convert('RGB'))
        t_pil = (time.perf_counter() - t0) / 30
        dev = coef.cuda()
        out = torch.empty((H, W, 3), dtype=torch.uint8, device='cuda')
        flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
        ms = []
        for it in range(25):
            flush.zero_()                                   # larger than L2: the stage reads its coefficients from HBM
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            J.reconstruct(info, dev, out=out)
            e1.record()
            torch.cuda.synchronize()
            if it >= 5:
                ms.append(e0.elapsed_time(e1))
        per_frame_us = {}
        same_b = True
        for n_batch in (8, 64):                             # 8: what the loops hand over at once (BACKBONE_LOOKAHEAD); 64: the kernels' own rate
            items = torch.stack([J.pack_item(data)] * n_batch).pin_memory()
            ms_b = []
            for it in range(25):
                flush.zero_()
                dev_items = items.cuda()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                frames = J.reconstruct_items(info, dev_items)
                e1.record()
                torch.cuda.synchronize()
                if it >= 5:
                    ms_b.append(e0.elapsed_time(e1))
            per_frame_us[n_batch] = float(np.median(ms_b)) * 1e3 / n_batch
            same_b = same_b and all(bool(torch.equal(f, out)) for f in frames)
        us_b = per_frame_us[64]
        same = same_b and bool(np.array_equal(out.cpu().numpy(), np.asarray(Image.open(io.BytesIO(data)).convert('RGB'))))
        us = float(np.median(ms)) * 1e3
        planes = sum(info.blocks_w[c] * info.blocks_h[c] * 64 for c in range(info.n_comp))
        alg = info.coef_count * 2 + 2 * planes + H * W * 3
        rec = {'frame': f'{H}x{W} quality {quality} subsampling {sub}', 'file_kb': round(len(data) / 1e3, 1), 'identical_to_pillow': same,
               'host_huffman_ms': round(t_huff * 1e3, 3), 'pillow_full_decode_ms': round(t_pil * 1e3, 3),
               'device_stage_us_one_frame_per_launch': round(us, 2), 'device_stage_us_per_frame_batch_of_8': round(per_frame_us[8], 2), 'device_stage_us_per_frame_batch_of_64': round(per_frame_us[64], 2),
               'algorithmic_bytes_per_frame': int(alg), 'achieved_gbs': round(alg / us_b / 1e3, 1),
               'hbm_peak_gbs': hbm, 'frac': round(alg / us_b / 1e3 / hbm, 4) if hbm else None,
               'l2': 'flushed between iterations (256 MB written)'}
        print(json.dumps(rec), flush=True)


if __name__ == '__main__':
    main()
