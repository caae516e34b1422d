"""Input pipeline of the training step: the reference's `load_data` + `read_labels`
(utils.py:38-121), `ImageData.image_processing` (utils.py:12-38) and the tf.data chain of BigGAN.build_model
(BigGAN.py:771-787: from_tensor_slices -> shuffle_and_repeat(dataset_num) -> map_and_batch(batch, 16 parallel, drop
remainder) -> prefetch_to_device(gpu, 4)).

Split by where the work belongs on a B200 box:
  * host: file listing, labels, an epoch-wise full shuffle that repeats for ever, PNG decode on a thread pool
    (`libbiggan_io.so`, C ABI `include/biggan_io.h`; ctypes releases the GIL) straight into pinned staging buffers;
  * device: ONE kernel per batch (`bg_image_prepare`) does the bilinear resize to `img_size`, the random left-right
    flip and `x / 127.5 - 1`, writing the layout the discriminator reads (bf16 mode: 16 zero-padded channels), on a side
    stream, `prefetch` batches ahead; the consumer only waits on an event.

`HostBatches` is the host half alone (uint8 batches; usable and tested without a GPU); `DevicePrefetcher` adds the
device half and refuses to run without CUDA -- there is no CPU substitute for it.
"""
import csv
import ctypes
import os
import threading
from concurrent.futures import ThreadPoolExecutor
from glob import glob

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_PKG, "libbiggan_io.so")
if not os.path.exists(_SO):
    raise RuntimeError("libbiggan_io.so not found at %s -- build it with `make -C biggan_tensorflow_b200/csrc`" % _SO)
io_lib = ctypes.CDLL(_SO)
io_lib.bgio_last_error.restype = ctypes.c_char_p
io_lib.bgio_png_info.argtypes = [ctypes.c_void_p, ctypes.c_size_t] + [ctypes.POINTER(ctypes.c_int)] * 4
io_lib.bgio_png_decode.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]


def png_info(data):
    """-> (width, height, channels, bit_depth) of PNG bytes."""
    w, h, c, d = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    buf = (ctypes.c_char * len(data)).from_buffer_copy(data)
    if io_lib.bgio_png_info(buf, len(data), w, h, c, d) != 0:
        raise ValueError(io_lib.bgio_last_error().decode())
    return w.value, h.value, c.value, d.value


def decode_png(data, channels=0, out=None):
    """tf.image.decode_png(data, channels): uint8 [h, w, channels] (0 = as stored).  `out`: optional preallocated array."""
    w, h, c, _ = png_info(data)
    oc = channels or c
    if out is None:
        out = np.empty((h, w, oc), dtype=np.uint8)
    elif out.shape != (h, w, oc) or out.dtype != np.uint8 or not out.flags.c_contiguous:
        raise ValueError("decode_png: `out` must be a contiguous uint8 array of shape %s" % ((h, w, oc),))
    buf = (ctypes.c_char * len(data)).from_buffer_copy(data)
    if io_lib.bgio_png_decode(buf, len(data), out.ctypes.data_as(ctypes.c_void_p), out.nbytes, channels) != 0:
        raise ValueError(io_lib.bgio_last_error().decode())
    return out


def read_labels(path):
    """utils.py:41-51: tab-separated `filename <tab> v0 <tab> v1 ...` -> {filename: [floats]}."""
    labels = {}
    with open(path, 'r') as f:
        for row in csv.reader(f, delimiter='\t', quotechar='"'):
            if row:
                labels[row[0]] = list(map(float, row[1:]))
    return labels


def load_data(dataset_name, label_file, weight_file=None, ignore_missing=False, n_labels=None, root="./dataset"):
    """utils.py:76-121 (custom datasets; the weight-file resampling is not carried over).  Returns (paths, labels or None)."""
    if weight_file:
        raise ValueError("weight_file resampling is not supported")
    x = sorted(glob(os.path.join(root, dataset_name, '*.*')))
    if not label_file:
        return x, None
    labels = read_labels(label_file)
    used = []
    for full in x:
        fn = os.path.basename(full)
        if fn not in labels:
            if not ignore_missing:
                raise RuntimeError("No label found for file " + fn)
            used.append([0.0] * n_labels)
        else:
            used.append(labels[fn])
    return x, used


class ImageData:
    """utils.py:12-38, host half: file name -> decoded uint8 pixels.  (Resize / flip / normalise: `bg_image_prepare`.)"""

    def __init__(self, load_size, channels, custom_dataset=True, flip=True):
        self.load_size, self.channels, self.custom_dataset, self.flip = load_size, channels, custom_dataset, flip

    def image_processing(self, filename, out=None):
        with open(filename, "rb") as f:
            return decode_png(f.read(), self.channels, out=out)


class HostBatches:
    """shuffle_and_repeat(len(data)) + map_and_batch(batch_size, drop_remainder=True): an endless iterator of
    (uint8 images [B, h, w, c], float32 labels [B, n] or None, flips int32 [B]).  All images of a batch must share
    one size (datasets are stored pre-scaled; the device resize brings them to img_size)."""

    def __init__(self, paths, labels, batch_size, channels=3, flip=True, seed=0, num_threads=16, alloc=None):
        if len(paths) < 1:
            raise ValueError("empty dataset")
        self.paths, self.labels = list(paths), None if labels is None else np.asarray(labels, dtype=np.float32)
        self.B, self.flip = int(batch_size), flip
        self.rng = np.random.Generator(np.random.PCG64(seed))
        self.loader = ImageData(0, channels, True, flip)
        self.pool = ThreadPoolExecutor(max_workers=num_threads)
        self.alloc = alloc or (lambda shape: np.empty(shape, dtype=np.uint8))
        self._order = iter(())

    def _next_index(self):
        while True:
            for i in self._order:
                return int(i)
            self._order = iter(self.rng.permutation(len(self.paths)))       # a fresh full shuffle every epoch

    def __iter__(self):
        return self

    def __next__(self):
        idx = [self._next_index() for _ in range(self.B)]
        with open(self.paths[idx[0]], "rb") as f:
            w, h, _, _ = png_info(f.read(64 * 1024))
        c = self.loader.channels
        batch = self.alloc((self.B, h, w, c))
        list(self.pool.map(lambda t: self.loader.image_processing(self.paths[t[1]], out=batch[t[0]]), enumerate(idx)))
        flips = (self.rng.random(self.B) < 0.5).astype(np.int32) if self.flip else np.zeros(self.B, np.int32)
        return batch, (None if self.labels is None else self.labels[idx]), flips


class DevicePrefetcher:
    """prefetch_to_device(gpu, n): keeps `prefetch` prepared batches in flight on a side stream.  Yields
    (images [B, S, S, Cd] in `dtype` on the GPU, labels [B, n] fp32 on the GPU or None)."""

    def __init__(self, host_batches, img_size, dtype=None, pad_channels=0, prefetch=4):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("DevicePrefetcher needs a CUDA device (there is no CPU path)")
        from . import _lib
        self.torch, self._lib = torch, _lib
        self.src, self.S = host_batches, int(img_size)
        self.dtype = dtype or torch.float32
        self.Cd = max(int(pad_channels), host_batches.loader.channels)
        self.stream = torch.cuda.Stream()
        self.depth = int(prefetch)
        self.queue = []
        host_batches.alloc = lambda shape: torch.empty(shape, dtype=torch.uint8).pin_memory().numpy()
        self._lock = threading.Lock()

    def _launch(self):
        torch = self.torch
        batch, labels, flips = next(self.src)
        B, h, w, c = batch.shape
        with torch.cuda.stream(self.stream):
            d_src = torch.from_numpy(batch).cuda(non_blocking=True)
            d_flip = torch.from_numpy(flips).cuda(non_blocking=True)
            dst = torch.empty((B, self.S, self.S, self.Cd), dtype=self.dtype, device="cuda")
            self._lib.call("bg_image_prepare", d_src, d_flip, dst, B, h, w, c, self.S, self.Cd,
                           self._lib.BG_BF16 if self.dtype == torch.bfloat16 else self._lib.BG_F32, self.stream.cuda_stream)
            d_lab = None if labels is None else torch.from_numpy(np.ascontiguousarray(labels)).cuda(non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.queue.append((ev, dst, d_lab, (batch, d_src, d_flip)))       # keep the staging buffers alive until consumed

    def __iter__(self):
        return self

    def __next__(self):
        with self._lock:
            while len(self.queue) < self.depth:
                self._launch()
            ev, dst, d_lab, _keep = self.queue.pop(0)
        cur = self.torch.cuda.current_stream()
        cur.wait_event(ev)
        dst.record_stream(cur)                 # allocated on the side stream, consumed on the caller's
        if d_lab is not None:
            d_lab.record_stream(cur)
        return dst, d_lab
