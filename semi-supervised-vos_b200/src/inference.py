"""`inference` command: same click surface and `inference_command_impl` signature as the
reference's src/inference.py:18-113."""
import os
from pathlib import Path

import click
import torch
import torch.utils.data
from loguru import logger

from src.config import Config
from src.model.vos_net import VOSNet
from src.utils.datasets import InferenceDataset
from src.utils.inference_utils import amp_enabled
from src.utils.inference_utils import (inference_2_scale, inference_3_scale, inference_hor_flip,
                                       inference_multimodel, inference_single, inference_ver_flip)
from src.utils.utils import load_model


@click.command(name='inference')
@click.option('--ref_num', '-n', type=int, default=9, help='Number of reference frames for inference.')
@click.option('--data', '-d', type=click.Path(file_okay=False, dir_okay=True), required=True,
              help='Path to inference dataset folder.')
@click.option('--resume', '-r', type=click.Path(file_okay=True, dir_okay=False), required=True,
              help='Path to the trained checkpoint.')
@click.option('--model', '-m', type=click.Choice(['resnet18', 'resnet50', 'resnet101', 'facebook']), default='resnet50',
              help='Network architecture, resnet18, resnet50, resnet101 or facebook.')
@click.option('--temperature', '-t', type=float, default=1.0, help='Temperature parameter.')
@click.option('--frame_range', type=int, default=40, help='Range of frames for inference.')
@click.option('--sigma_1', type=float, default=8.0, help='Smaller sigma in the motion model for dense spatial weight')
@click.option('--sigma_2', type=float, default=21.0, help='Larger sigma in the motion model for dense spatial weight.')
@click.option('--save', '-s', type=click.Path(file_okay=False, dir_okay=True), required=True,
              help='Path to save predictions.')
@click.option('--device', type=click.Choice(['cpu', 'cuda']), default='cuda', help='Device to run computing on.')
@click.option('--inference-strategy',
              type=click.Choice(['single', 'hor-flip', 'vert-flip', '2-scale', 'multimodel', 'hor-2-scale', '3-scale']),
              default='single', help='Inference strategy.')
@click.option('--additional-model', type=click.Path(file_okay=True, dir_okay=False), required=False,
              help='Path to the additional checkpoint.')
@click.option('--additional-model-type', type=click.STRING, required=False, default='resnet50',
              help='Type of additional model type.')
@click.option('--probability/--no-probability', default=False, required=False,
              help='Should probability or labels be propagated.')
@click.option('--scale', default=1.15, required=False, type=click.FLOAT, help='Scale for 2nd image in 2-scale strategy.')
@click.option('--fusion', default='mean', type=click.Choice(['maximum', 'minimum', 'mean']),
              help='Fusion operation for probability propagation.')
def inference_command(ref_num, data, resume, model, temperature, frame_range, sigma_1, sigma_2, save, device,
                      inference_strategy, additional_model, additional_model_type, probability, scale, fusion):
    inference_command_impl(ref_num, data, resume, model, temperature, frame_range, sigma_1, sigma_2, save, device,
                           inference_strategy, additional_model, additional_model_type, probability, scale, fusion)


def _load_net(arch, checkpoint):
    """Checkpoint -> eval-mode network on Config.DEVICE.  The ImageNet initialisation the reference downloads first
    (vos_net.py:17-19) is overwritten by the checkpoint anyway and is skipped.  ResNet trunks run in their cuDNN-fused
    inference form (BatchNorm folded, bias / ReLU / residual in the convolution epilogue, fp16 as under the reference's
    autocast); VOS_FUSE_BACKBONE=0 keeps the plain module."""
    net = load_model(VOSNet(model=arch, pretrained=False), checkpoint).to(Config.DEVICE).eval()
    if not amp_enabled():          # parity mode against the reference's fp32 CPU path: plain module, true fp32 convolutions
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        return net
    if Config.DEVICE.type == 'cuda' and arch in ('resnet18', 'resnet50', 'resnet101') and os.environ.get('VOS_FUSE_BACKBONE', '1') != '0':
        from vosb200.fused_backbone import FusedVOSNet
        return FusedVOSNet(net)
    return net


def inference_command_impl(ref_num, data, resume, model, temperature, frame_range, sigma_1, sigma_2, save, device,
                           inference_strategy, additional_resume, additional_model_type, probability_propagation,
                           scale, reduction, disable=False):
    if Config.DEVICE.type != device:
        Config.DEVICE = torch.device(device)
    # `torchrun --nproc-per-node N main.py inference ...`: one process per GPU, whole videos sharded over the ranks (the
    # reference resets all state at a video boundary, inference_utils.py:28-48, so the masks do not depend on the sharding);
    # every rank writes the PNGs of its own videos, no collective anywhere
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    if world > 1 and device == 'cuda':
        Config.DEVICE = torch.device('cuda', int(os.environ.get('LOCAL_RANK', rank)))
        torch.cuda.set_device(Config.DEVICE)
    model = _load_net(model, resume)
    additional_model = _load_net(additional_model_type, additional_resume) if inference_strategy == 'multimodel' else None

    videos = None
    if world > 1:
        from vosb200.shard import assign_lpt, sequence_cost
        counts = sorted((d.name, sum(1 for _ in d.iterdir())) for d in (Path(data) / 'JPEGImages/480p').iterdir() if d.is_dir())
        mine = assign_lpt([sequence_cost(n, 1, ref_num) for _, n in counts], world)[rank]
        videos = {counts[i][0] for i in mine}
        logger.info(f'rank {rank}/{world}: {len(videos)} of {len(counts)} videos')
        if not videos:
            return
    dataset = InferenceDataset(str(Path(data) / 'JPEGImages/480p'), disable=disable,
                               inference_strategy=inference_strategy, scale=scale,
                               raw='coef' if os.environ.get('VOS_GPU_JPEG', '1') != '0' else True, videos=videos)
    # the reference decodes with one worker (inference.py:75-78); JPEG decode is the slowest stage once propagation runs on
    # the GPU, so it is spread over workers here (same PIL decode, same order: shuffle=False)
    cpus = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    # (An in-process loader on the library's own threads -- vosjpeg_decode_files_host for 8 files at a time, one group ahead -- was
    # measured against the worker processes: 540-598 vs 684-719 frames/s on 16 host threads, 450-475 vs 519-535 on 4: the producer
    # thread shares the interpreter with the launch loop.  Workers stay.)
    loader = torch.utils.data.DataLoader(dataset, batch_size=1, shuffle=False, num_workers=max(1, min(12, cpus - 4, cpus)) if cpus > 8 else min(8, cpus),
                                         pin_memory=True, prefetch_factor=4, persistent_workers=False)
    annotation_dir = Path(data) / 'Annotations/480p'
    last_video = sorted(v.name for v in annotation_dir.glob('*') if videos is None or v.name in videos)[0]
    if probability_propagation:
        # probability propagation keeps a dense label record per reference pixel: at most MAX_DENSE_CLASSES classes.  Checked for
        # every video BEFORE anything runs, so a run does not abort half-way with earlier videos already written (the
        # reference itself takes any d = max(label) + 1, predict.py:113)
        import numpy as np
        from PIL import Image
        from vosb200._capi import MAX_DENSE_CLASSES
        for v in sorted(annotation_dir.glob('*')):
            if (videos is None or v.name in videos) and (v / '00000.png').is_file():
                d = int(np.array(Image.open(v / '00000.png')).max()) + 1
                if d > MAX_DENSE_CLASSES:
                    raise ValueError(f'{v.name}: {d} classes in the first annotation; --probability-propagation supports at most '
                                     f'{MAX_DENSE_CLASSES} (index-label propagation, the default, takes up to 24)')
    common = (loader, len(dataset), annotation_dir, last_video, save, sigma_1, sigma_2, frame_range, ref_num,
              temperature, probability_propagation)
    with torch.no_grad():
        if inference_strategy == 'single':
            inference_single(model, *common, disable)
        elif inference_strategy == 'hor-flip':
            inference_hor_flip(model, *common, reduction, disable)
        elif inference_strategy == 'vert-flip':
            inference_ver_flip(model, *common, reduction, disable)
        elif inference_strategy in ('2-scale', 'hor-2-scale'):
            inference_2_scale(model, *common, scale, reduction, inference_strategy == 'hor-2-scale', disable)
        elif inference_strategy == 'multimodel':
            inference_multimodel(model, additional_model, *common, reduction, disable)
        elif inference_strategy == '3-scale':
            inference_3_scale(model, *common, scale, disable)
    logger.info('Inference done.')
