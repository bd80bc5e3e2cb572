"""The datasets either side of the propagation path (reference src/utils/datasets.py).
InferenceDataset (datasets.py:111-167): one frame per item, `(normalised CHW tensor, video_name)`; frames are grouped
by sub-directory.  TrainDataset (datasets.py:19-109): clips of `frame_num` consecutive frames of one video, all cropped
and flipped the same way, with their RGB annotations -- what `validation` feeds to the loss."""
from io import BytesIO
from pathlib import Path

import numpy as np
import torch
from loguru import logger
from PIL import Image, ImageOps
from torchvision import datasets, transforms
from torchvision.datasets.folder import make_dataset
from tqdm import tqdm

from src.utils.transforms import crop, get_crop_params

_NORMALIZE = dict(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])


class TrainDataset(datasets.ImageFolder):
    """Item = (frames (T,3,c,c) fp32 normalised, annotations (T,3,c,c) fp32 RGB 0..255, video index).  RNG draws per
    item, in the reference's order (datasets.py:72-86): horizontal flip, vertical flip, then the crop origin."""

    def __init__(self, img_root, annotation_root, cropping=256, frame_num=10, transform=None, target_transform=None,
                 color_jitter=False):
        super().__init__(img_root, transform=transform, target_transform=target_transform)
        if color_jitter:
            raise NotImplementedError('colour jitter is a training-time augmentation; validation runs without it')
        self.annotations = make_dataset(annotation_root, self.class_to_idx, extensions=('png', 'jpg', 'jpeg'))
        self.cropping, self.frame_num, self.color_jitter = cropping, frame_num, color_jitter
        self.rgb_normalize = transforms.Compose([transforms.ToTensor(), transforms.Normalize(**_NORMALIZE)])
        logger.info(f'Loading {len(self.imgs)} train images and annotations.')
        self.img_bytes = [Path(p).read_bytes() for p, _ in tqdm(self.imgs)]
        self.annotation_bytes = [Path(p).read_bytes() for p, _ in tqdm(self.annotations)]

    def _one_video(self, index):
        return self.imgs[index][1] == self.imgs[index + self.frame_num - 1][1]

    def __getitem__(self, index):
        T = self.frame_num
        index = min(index, len(self.imgs) - T)      # clips never run past the end of the dataset ...
        while not self._one_video(index):           # ... nor across a video boundary
            index -= 1
        h_flip = torch.rand(size=(1,)).item() < 0.5
        v_flip = torch.rand(size=(1,)).item() < 0.5
        frames, annotations, window = [], [], None
        for i in range(T):
            img = Image.open(BytesIO(self.img_bytes[index + i])).convert('RGB')
            ann = Image.open(BytesIO(self.annotation_bytes[index + i])).convert('RGB')
            if h_flip:
                img, ann = img.transpose(Image.FLIP_LEFT_RIGHT), ann.transpose(Image.FLIP_LEFT_RIGHT)
            if v_flip:
                img, ann = img.transpose(Image.FLIP_TOP_BOTTOM), ann.transpose(Image.FLIP_TOP_BOTTOM)
            if window is None:
                window = get_crop_params(img.size, self.cropping)
            frames.append(self.rgb_normalize(crop(img, *window)).numpy())
            annotations.append(np.asarray(crop(ann, *window)).transpose(2, 0, 1))
        return (torch.from_numpy(np.asarray(frames)).float(), torch.from_numpy(np.asarray(annotations)).float(),
                self.imgs[index + T - 1][1])


class InferenceDataset(datasets.ImageFolder):
    """`raw=True`: items carry the decoded frame as a uint8 (H,W,3) tensor instead of the normalised fp32 (3,H,W) one; the loops
    normalise it on the GPU (vosprop_normalize_u8: same arithmetic, same bits) -- ToTensor + Normalize cost more host time per
    480p frame than the JPEG decode.  `raw='coef'` (what `inference_command_impl` asks for unless VOS_GPU_JPEG=0): items of the
    strategies that feed the network the frame as decoded carry the frame's quantised DCT coefficients (int16, vosb200/jpeg.py)
    and the GPU finishes the decode with Pillow's exact pixels; other strategies and other JPEG flavours behave like raw=True."""

    def __init__(self, root, transform=None, target_transform=None, disable=False,
                 inference_strategy='single', scale=None, raw=False, videos=None):
        super().__init__(root, transform=transform, target_transform=target_transform)
        if videos is not None:         # this rank's share of the videos (src/inference.py under torchrun)
            keep = {self.class_to_idx[v] for v in videos}
            self.samples = [s for s in self.samples if s[1] in keep]
            self.imgs = self.samples
            self.targets = [s[1] for s in self.samples]
        self.raw = raw
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
        if self.raw == 'coef' and self.inference_strategy in ('single', 'multimodel', '3-scale'):
            # JPEG front end of libvosprop (vosb200/jpeg.py): the worker does the Huffman half of the decode and ships the
            # quantised DCT coefficients; de-quantisation, inverse DCT, chroma up-sampling and colour conversion run on the GPU
            # with Pillow's exact pixels (inference_utils._to_device).  Files outside that path keep Pillow below.
            from vosb200 import jpeg
            try:
                return jpeg.pack_item(self.img_bytes[index]), self.idx_to_class[video_index]
            except jpeg.Unsupported:
                pass
        img = Image.open(BytesIO(self.img_bytes[index]))
        if img.mode != 'RGB':                      # convert() copies even when there is nothing to convert (1.6 ms at 480p)
            img = img.convert('RGB')
        # the reference resizes to ceil(size) == size with ANTIALIAS: an identity resample
        tensor = (lambda im: torch.from_numpy(np.array(im))) if self.raw else (lambda im: self.rgb_normalize(np.asarray(im)))
        normalized = tensor(img)
        video = self.idx_to_class[video_index]
        if self.inference_strategy == 'hor-flip':
            return (normalized, tensor(ImageOps.mirror(img))), video
        if self.inference_strategy == 'vert-flip':
            return (normalized, tensor(ImageOps.flip(img))), video
        if self.inference_strategy in ('2-scale', 'hor-2-scale'):
            size2 = tuple(int(v) for v in np.ceil(np.array(img.size) * self.scale))
            src = ImageOps.mirror(img) if self.inference_strategy == 'hor-2-scale' else img
            return (normalized, tensor(src.resize(size2, Image.LANCZOS))), video
        return normalized, video

    def __len__(self):
        return len(self.imgs)
