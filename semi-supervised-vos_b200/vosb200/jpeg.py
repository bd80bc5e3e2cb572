"""JPEG front end over include/vos_jpeg.h: Huffman decoding on the host (any thread, GIL released), reconstruction on the GPU,
pixels bit-identical to Pillow's `Image.open(...).convert('RGB')` (the reference's loader, src/utils/datasets.py:141-143) for
baseline / extended-sequential JPEGs with 4:4:4, 4:2:2 or 4:2:0 sampling (or grey); `Unsupported` for anything else, so that the
caller keeps Pillow for that file.

    info = parse(data)                       # header
    coef = entropy_decode(data, info)        # (coef_count,) int16, pinned when a GPU is there
    rgb = reconstruct(info, coef.cuda())     # (H, W, 3) uint8 on the device
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _capi as capi

ERR_UNSUPPORTED = -2


class Info(C.Structure):
    _fields_ = [('width', C.c_int32), ('height', C.c_int32), ('n_comp', C.c_int32),
                ('h_samp', C.c_int32 * 3), ('v_samp', C.c_int32 * 3), ('blocks_w', C.c_int32 * 3), ('blocks_h', C.c_int32 * 3),
                ('coef_offset', C.c_int64 * 3), ('coef_count', C.c_int64), ('scan_offset', C.c_int64),
                ('restart_interval', C.c_int32), ('dc_table', C.c_int32 * 3), ('ac_table', C.c_int32 * 3),
                ('quant', (C.c_uint16 * 64) * 3)]


class JpegError(RuntimeError):
    pass


class Unsupported(JpegError):
    """A JPEG flavour outside the bit-exact path (progressive, CMYK, unusual sampling ...): decode it with Pillow."""


EXPORTS = {
    'vosjpeg_last_error': (C.c_char_p, []),
    'vosjpeg_parse': (C.c_int, [C.c_char_p, C.c_int64, C.POINTER(Info)]),
    'vosjpeg_entropy_decode': (C.c_int, [C.c_char_p, C.c_int64, C.POINTER(Info), C.c_void_p]),
    'vosjpeg_decode_files_host': (C.c_int, [C.POINTER(C.c_char_p), C.POINTER(C.c_int64), C.c_int32, C.POINTER(C.c_void_p), C.c_int64, C.c_int32,
                                            C.c_int32, C.POINTER(C.c_int32)]),
    'vosjpeg_scratch_bytes': (C.c_int64, [C.POINTER(Info)]),
    'vosjpeg_reconstruct': (C.c_int, [C.POINTER(Info), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'vosjpeg_reconstruct_batch': (C.c_int, [C.POINTER(Info), C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                            C.c_void_p]),
    'vosjpeg_reconstruct_host': (C.c_int, [C.POINTER(Info), C.c_void_p, C.c_void_p]),
}
_bound = None


def _lib() -> C.CDLL:
    global _bound
    if _bound is None:
        lib = capi.lib()
        for name, (res, args) in EXPORTS.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _bound = lib
    return _bound


def _check(rc: int):
    if rc < 0:
        msg = _lib().vosjpeg_last_error().decode()
        raise (Unsupported if rc == ERR_UNSUPPORTED else JpegError)(msg)


def parse(data: bytes) -> Info:
    info = Info()
    _check(_lib().vosjpeg_parse(data, len(data), C.byref(info)))
    return info


def entropy_decode(data: bytes, info: Info, out: Optional[torch.Tensor] = None, pinned: Optional[bool] = None) -> torch.Tensor:
    """Quantised DCT coefficients of the whole scan as a flat int16 tensor on the host (layout: vos_jpeg.h)."""
    if out is None:
        pin = torch.cuda.is_available() if pinned is None else pinned
        out = torch.empty(info.coef_count, dtype=torch.int16, pin_memory=pin)
    if out.dtype != torch.int16 or out.numel() != info.coef_count or out.is_cuda or not out.is_contiguous():
        raise ValueError('out: contiguous host int16 tensor of info.coef_count elements')
    _check(_lib().vosjpeg_entropy_decode(data, len(data), C.byref(info), out.data_ptr()))
    return out


def reconstruct(info: Info, coef: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Coefficients -> (H, W, 3) uint8 RGB.  Device tensor: two kernels on the current stream; host tensor: the same arithmetic
    on one CPU thread."""
    if coef.dtype != torch.int16 or coef.numel() != info.coef_count or not coef.is_contiguous():
        raise ValueError('coef: contiguous int16 tensor of info.coef_count elements')
    if out is None:
        out = torch.empty((info.height, info.width, 3), dtype=torch.uint8, device=coef.device)
    if out.dtype != torch.uint8 or tuple(out.shape) != (info.height, info.width, 3) or out.device != coef.device or not out.is_contiguous():
        raise ValueError('out: contiguous (H, W, 3) uint8 tensor on the coefficients\' device')
    if not coef.is_cuda:
        _check(_lib().vosjpeg_reconstruct_host(C.byref(info), coef.data_ptr(), out.data_ptr()))
        return out
    with torch.cuda.device(coef.device):
        scratch = torch.empty(_lib().vosjpeg_scratch_bytes(C.byref(info)), dtype=torch.uint8, device=coef.device)
        stream = torch.cuda.current_stream(coef.device)
        _check(_lib().vosjpeg_reconstruct(C.byref(info), coef.data_ptr(), scratch.data_ptr(), out.data_ptr(), C.c_void_p(stream.cuda_stream)))
        scratch.record_stream(stream)
    return out


def decode(data: bytes, device=None) -> torch.Tensor:
    """bytes of a JPEG file -> (H, W, 3) uint8 RGB on `device` (None: the host)."""
    info = parse(data)
    coef = entropy_decode(data, info, pinned=device is not None)
    if device is not None:
        coef = coef.to(device, non_blocking=True)
    return reconstruct(info, coef)


# ---- loader items: one flat int16 tensor per frame = [Info struct, padded to 16 bytes][coefficients] ----------------------------
_HDR_I16 = (C.sizeof(Info) + 15) // 16 * 8          # header length in int16 elements (coefficients stay 16-byte aligned)


def pack_item(data: bytes) -> torch.Tensor:
    """Host half of the decode for a DataLoader worker (no CUDA): header + Huffman-decoded coefficients in one int16 tensor."""
    info = parse(data)
    item = torch.empty(_HDR_I16 + info.coef_count, dtype=torch.int16)
    hdr = bytes(info) + b'\0' * (_HDR_I16 * 2 - C.sizeof(Info))
    item[:_HDR_I16] = torch.frombuffer(bytearray(hdr), dtype=torch.int16)
    entropy_decode(data, info, out=item[_HDR_I16:])
    return item


def unpack_items(items: torch.Tensor, device) -> torch.Tensor:
    """(n, L) int16 items of equal geometry (host) -> (n, H, W, 3) uint8 RGB frames on `device`."""
    if items.dtype != torch.int16 or items.dim() != 2:
        raise ValueError('items: (n, L) int16')
    infos = [Info.from_buffer_copy(items[i, :_HDR_I16].contiguous().numpy().tobytes()[:C.sizeof(Info)]) for i in range(items.shape[0])]
    for info in infos:
        same = all(list(getattr(info, k)) == list(getattr(infos[0], k)) for k in ('h_samp', 'v_samp', 'blocks_w', 'blocks_h'))
        if (info.height, info.width, info.n_comp) != (infos[0].height, infos[0].width, infos[0].n_comp) or not same or \
                _HDR_I16 + info.coef_count != items.shape[1]:
            raise ValueError('items of one batch must share their geometry')
    dev = items.to(device, non_blocking=True)
    if not dev.is_cuda:
        out = torch.empty((items.shape[0], infos[0].height, infos[0].width, 3), dtype=torch.uint8)
        for i, inf in enumerate(infos):
            reconstruct(inf, dev[i, _HDR_I16:], out=out[i])
        return out
    return reconstruct_items(infos[0], dev)


def reconstruct_items(info: Info, dev: torch.Tensor) -> torch.Tensor:
    """Device stage for a batch of loader items already on the GPU ((n, L) int16, one geometry = `info`'s): one launch pair; each
    frame's quantisation tables are read from its own header, which travelled with it."""
    n = dev.shape[0]
    out = torch.empty((n, info.height, info.width, 3), dtype=torch.uint8, device=dev.device)
    quant_off = Info.quant.offset // 2
    with torch.cuda.device(dev.device):
        per_frame = (_lib().vosjpeg_scratch_bytes(C.byref(info)) + 7) // 8 * 8
        scratch = torch.empty(n * per_frame, dtype=torch.uint8, device=dev.device)
        stream = torch.cuda.current_stream(dev.device)
        _check(_lib().vosjpeg_reconstruct_batch(C.byref(info), n, dev.data_ptr() + 2 * _HDR_I16, dev.shape[1], dev.data_ptr() + 2 * quant_off,
                                                dev.shape[1], scratch.data_ptr(), out.data_ptr(), C.c_void_p(stream.cuda_stream)))
        scratch.record_stream(stream)
        dev.record_stream(stream)
    return out


def pack_items_threaded(datas, capacity: int, threads: int, pinned: bool = False):
    """Host half of the decode for a batch of files on the library's own threads (GIL released for the whole batch): returns
    (buffer (n, capacity) int16, status list).  Row i holds file i's loader item in its first _HDR_I16 + coef_count values when
    status[i] == 0; status ERR_UNSUPPORTED: decode that file with Pillow."""
    n = len(datas)
    buf = torch.empty((n, capacity), dtype=torch.int16, pin_memory=pinned)
    arr = (C.c_char_p * n)(*datas)
    sizes = (C.c_int64 * n)(*[len(d) for d in datas])
    ptrs = (C.c_void_p * n)(*[buf.data_ptr() + 2 * capacity * i for i in range(n)])
    status = (C.c_int32 * n)()
    _check(_lib().vosjpeg_decode_files_host(arr, sizes, n, ptrs, capacity, _HDR_I16, threads, status))
    return buf, list(status)


def item_length(data: bytes) -> int:
    """Values a loader item of this file takes (header + coefficients)."""
    return _HDR_I16 + parse(data).coef_count


def item_values(item: torch.Tensor) -> int:
    """Length (in int16 values) of the loader item that starts at item[0] (its header says how many coefficients follow)."""
    info = Info.from_buffer_copy(item[:_HDR_I16].contiguous().numpy().tobytes()[:C.sizeof(Info)])
    return _HDR_I16 + info.coef_count
