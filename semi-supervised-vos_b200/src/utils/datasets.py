"""InferenceDataset: the caller side of the hot path (reference src/utils/datasets.py:111-167).
One frame per item, `(normalised CHW tensor, video_name)`; frames are grouped by sub-directory."""
from io import BytesIO
from pathlib import Path

import numpy as np
from loguru import logger
from PIL import Image, ImageOps
from torchvision import datasets, transforms
from tqdm import tqdm


class InferenceDataset(datasets.ImageFolder):
    def __init__(self, root, transform=None, target_transform=None, disable=False,
                 inference_strategy='single', scale=None):
        super().__init__(root, transform=transform, target_transform=target_transform)
        self.rgb_normalize = transforms.Compose([
            transforms.ToTensor(),
            transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
        logger.info(f'Loading {len(self.imgs)} inference images.')
        self.img_bytes = [Path(path).read_bytes() for path, _ in tqdm(self.imgs, disable=disable)]
        logger.info(f'Loaded {len(self.img_bytes)} inference images.')
        self.idx_to_class = {v: k for k, v in self.class_to_idx.items()}
        self.inference_strategy = inference_strategy
        self.scale = scale

    def __getitem__(self, index):
        _, video_index = self.imgs[index]
        img = Image.open(BytesIO(self.img_bytes[index])).convert('RGB')
        # the reference resizes to ceil(size) == size with ANTIALIAS: an identity resample
        normalized = self.rgb_normalize(np.asarray(img))
        video = self.idx_to_class[video_index]
        if self.inference_strategy == 'hor-flip':
            return (normalized, self.rgb_normalize(np.asarray(ImageOps.mirror(img)))), video
        if self.inference_strategy == 'vert-flip':
            return (normalized, self.rgb_normalize(np.asarray(ImageOps.flip(img)))), video
        if self.inference_strategy in ('2-scale', 'hor-2-scale'):
            size2 = tuple(int(v) for v in np.ceil(np.array(img.size) * self.scale))
            src = ImageOps.mirror(img) if self.inference_strategy == 'hor-2-scale' else img
            return (normalized, self.rgb_normalize(np.asarray(src.resize(size2, Image.LANCZOS)))), video
        return normalized, video

    def __len__(self):
        return len(self.imgs)
