"""Device time of the validation criterion on one GPU: the engine-backed CrossEntropy.forward of
src/model/loss.py on a batch of 16 clips (9 reference frames + 1 target of 32 x 32 embeddings, 22 classes -- what
src/validation.py feeds it with --bs 16), next to the reference's own formulation (loss.py:13-66: bmm -> softmax over
a (B, 9216, 1024) tensor -> bmm -> log -> NLL) executed by torch on the same GPU.  One JSON object per line."""
import json
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
for p in (str(REPO), str(REPO / 'semi-supervised-vos_b200')):
    sys.path.insert(0, p)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from src.model.loss import CrossEntropy  # noqa: E402


def torch_formulation(ref, target, ref_cls, target_cls, d, temperature):
    B, R, K, H, W = ref.shape
    sim = ref.permute(0, 1, 3, 4, 2).reshape(B, -1, K).bmm(target.reshape(B, K, -1)) * temperature
    sim = sim.softmax(dim=1)
    onehot = torch.zeros(B, R, d, H, W, device=ref.device, dtype=ref.dtype).scatter_(2, ref_cls.unsqueeze(2), 1)
    pred = onehot.transpose(1, 2).reshape(B, d, -1).bmm(sim).reshape(B, d, H, W)
    return F.nll_loss(torch.log(pred.float() + 1e-14), target_cls)


def timed(fn, reps=20, warm=5):
    for _ in range(warm):
        out = fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        out = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, float(out)


def main():
    dev = torch.device('cuda', 0)
    B, R, K, H, W, d = 16, 9, 256, 32, 32, 22
    g = torch.Generator(device=dev).manual_seed(0)
    cls = torch.randint(0, d, (B, R + 1, H // 4, W // 4), device=dev, generator=g).repeat_interleave(4, 2).repeat_interleave(4, 3)
    proto = torch.randn(d, K, device=dev, generator=g)
    feats = (proto[cls].permute(0, 1, 4, 2, 3) + 1.5 * torch.randn(B, R + 1, K, H, W, device=dev, generator=g)) * 0.075
    crit = CrossEntropy(1.0)
    for dtype in (torch.float32, torch.float16):
        f = feats.to(dtype)
        ref, target, rc, tc = f[:, :-1], f[:, -1], cls[:, :-1], cls[:, -1]
        ms_e, loss_e = timed(lambda: crit(ref, target, rc, tc))
        ms_t, loss_t = timed(lambda: torch_formulation(ref, target, rc, tc, d, 1.0))
        print(json.dumps({'config': 'validation criterion', 'batch_clips': B, 'refs': R, 'pixels': H * W, 'classes': d,
                          'embedding_dtype': str(dtype).replace('torch.', ''), 'engine_ms_per_batch': round(ms_e, 3),
                          'torch_formulation_ms_per_batch': round(ms_t, 3), 'engine_loss': loss_e, 'torch_loss': loss_t,
                          'clips_per_s_engine': round(B / ms_e * 1e3, 1)}), flush=True)


if __name__ == '__main__':
    main()
