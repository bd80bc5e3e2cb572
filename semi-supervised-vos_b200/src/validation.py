"""`validation` command: scores every checkpoint of a directory with the training criterion on random 256x256 clips
(reference src/validation.py:29-99).  Same options; `--loss cross_entropy` runs on the propagation engine, the other
criteria (focal / contrastive / triplet + miners) are not built."""
import json
import math
from pathlib import Path

import click
import numpy as np
import torch
import torch.utils.data
from loguru import logger
from tqdm import tqdm

from src.config import Config
from src.model.loss import CrossEntropy
from src.model.vos_net import VOSNet
from src.train import step
from src.utils.datasets import TrainDataset
from src.utils.utils import annotation_centroids, load_model

_MINERS = ['default', 'kernel_7x7', 'temporal', 'one_back_one_ahead', 'euclidean', 'manhattan', 'chebyshev', 'skeleton',
           'skeleton_nearest_negative', 'skeleton_temporal']


@click.command(name='validation')
@click.option('--data', '-d', type=click.Path(file_okay=False, dir_okay=True), required=True, help='Path to dataset.')
@click.option('--checkpoints', '-c', type=click.Path(dir_okay=True, file_okay=False), help='Path to checkpoints.')
@click.option('--bs', type=int, default=16, help='Batch size.')
@click.option('--loss', type=click.Choice(['cross_entropy', 'focal', 'contrastive', 'triplet']), default='cross_entropy',
              help='Loss function to use.')
@click.option('--miner', type=click.Choice(_MINERS), default='default', help='Triplet loss miner.')
@click.option('--margin', type=click.FloatRange(min=0.0, max=1.0), default=0.1, help='Triplet loss margin.')
@click.option('--loss_weight', type=click.FloatRange(min=0.0), default=6.0, help='Weight of triplet loss.')
@click.option('--output', '-o', type=click.Path(dir_okay=False, file_okay=True), help='Path to output JSON.')
@click.option('--model', '-m', 'arch', type=click.Choice(['resnet18', 'resnet50', 'resnet101']), default='resnet50',
              help='Network architecture of the checkpoints (the reference hard-codes resnet50).')
@click.option('--workers', type=int, default=8, help='DataLoader workers (the reference hard-codes 8).')
def validation_command(data, checkpoints, bs, loss, miner, margin, loss_weight, output, arch, workers):
    validation_command_impl(data, checkpoints, bs, loss, output, arch, workers)


def validation_command_impl(data, checkpoints, bs, loss, output, arch='resnet50', workers=8, temperature=1.0):
    logger.info('Validation started.')
    if loss != 'cross_entropy':
        raise click.ClickException(f'--loss {loss} is not built: only cross_entropy runs on the propagation engine')
    if Config.DEVICE.type != 'cuda':
        raise click.ClickException('validation runs on the CUDA propagation engine; no CUDA device is visible')
    criterion = CrossEntropy(temperature=temperature).to(Config.DEVICE)
    dataset = TrainDataset(Path(data) / 'JPEGImages/480p', Path(data) / 'Annotations/480p', frame_num=10, color_jitter=False)
    loader = torch.utils.data.DataLoader(dataset, batch_size=bs, shuffle=False, pin_memory=True, num_workers=workers,
                                         drop_last=True)
    batches = math.ceil(len(dataset) / bs)
    local = Path('./annotation_centroids.npy')
    centroids = np.load(local) if local.is_file() else annotation_centroids()
    centroids = torch.Tensor(centroids).float().to(Config.DEVICE)

    losses = {}
    for checkpoint in tqdm(sorted(Path(checkpoints).glob('*.pth.tar')), desc='Validating checkpoints: '):
        # load_model retries through nn.DataParallel for 'module.'-prefixed checkpoints (validation.py:88-92)
        model = load_model(VOSNet(model=arch, pretrained=False).to(Config.DEVICE), str(checkpoint.absolute()))
        losses[checkpoint.name] = step(loader, model.eval(), criterion, None, 0, centroids, batches, mode='val')
    with Path(output).open(mode='w') as writer:
        json.dump(losses, writer)
    logger.info('Validation finished.')
    return losses
