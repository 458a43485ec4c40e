"""The PyTorch custom-op shim: ``torch.ops.spb200.forward`` / ``torch.ops.spb200.detect`` / ``torch.ops.spb200.detect_u8``.

Thin wrappers over the same C ABI as everything else (include/spb200.h): the dispatcher sees ordinary ops (CUDA tensors in,
CUDA tensors out, work enqueued on torch's current stream), so the drop-in ``spb200.SuperPoint`` module composes with the
rest of PyTorch - streams, ``torch.no_grad``, ``torch.compile`` / FX tracing through the registered fake kernels - without
PyTorch ever computing anything of the path itself.  An engine is addressed by an integer id (``register(engine)``):
custom ops take tensors and scalars only.

    prob_map, desc_map, logits = torch.ops.spb200.forward(image, engine_id)         # SuperPoint.forward, superpoint.py:91-115
    count, xy, conf, desc = torch.ops.spb200.detect(image, engine_id, capacity)     # InferenceWrapper.run, inferencewrapper.py:29-46
"""
import threading

import torch

_engines = {}
_lock = threading.Lock()
_next_id = [1]


def register(engine):
    """Make `engine` (spb200.Engine) addressable from the ops; returns its id (stable for the engine's lifetime)."""
    with _lock:
        eid = getattr(engine, '_op_id', None)
        if eid is None:
            eid = _next_id[0]
            _next_id[0] += 1
            engine._op_id = eid
        _engines[eid] = engine
        return eid


def unregister(engine):
    with _lock:
        _engines.pop(getattr(engine, '_op_id', None), None)


def _engine(eid):
    try:
        return _engines[int(eid)]
    except KeyError:
        raise RuntimeError('spb200 op: no engine registered under id %d (spb200.ops.register)' % int(eid))


def _desc_dtype(eid):
    """fp32 unless the engine was switched to 16-bit descriptors (spb200_set_descriptor_format); fp32 for an unknown id (tracing)."""
    e = _engines.get(int(eid))
    return getattr(e, 'desc_dtype', torch.float32) if e is not None else torch.float32


@torch.library.custom_op('spb200::forward', mutates_args=(), device_types='cuda')
def forward(image: torch.Tensor, engine: int) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    return _engine(engine).forward(image)


@forward.register_fake
def _(image, engine):
    b, _, h, w = image.shape
    return (image.new_empty((b, h, w)), image.new_empty((b, 128, h // 8, w // 8)), image.new_empty((b, 65, h // 8, w // 8)))


@torch.library.custom_op('spb200::detect', mutates_args=(), device_types='cuda')
def detect(image: torch.Tensor, engine: int, capacity: int) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    count, xy, conf, desc, _ = _engine(engine).detect(image, capacity)
    return count, xy, conf, desc


@detect.register_fake
def _(image, engine, capacity):
    b = image.shape[0]
    return (image.new_empty((b,), dtype=torch.int32), image.new_empty((b, capacity, 2), dtype=torch.int32),
            image.new_empty((b, capacity), dtype=torch.float32), image.new_empty((b, capacity, 128), dtype=_desc_dtype(engine)))


@torch.library.custom_op('spb200::detect_u8', mutates_args=(), device_types='cuda')
def detect_u8(frames: torch.Tensor, engine: int, capacity: int) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    count, xy, conf, desc, _ = _engine(engine).detect_u8(frames, capacity)
    return count, xy, conf, desc


@detect_u8.register_fake
def _(frames, engine, capacity):
    b = frames.shape[0]
    return (frames.new_empty((b,), dtype=torch.int32), frames.new_empty((b, capacity, 2), dtype=torch.int32),
            frames.new_empty((b, capacity), dtype=torch.float32), frames.new_empty((b, capacity, 128), dtype=_desc_dtype(engine)))
