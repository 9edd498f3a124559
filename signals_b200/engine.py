"""The GPU evaluator behind ``Emitter.respond`` / ``BoundPort.request``.

``Engine.render(emitter, loc)`` is what the reference's recursion
(/root/reference/src/signals/chain/__init__.py:296-300) becomes: compile the sub-graph once
(``signals_b200.plan.lower`` -> ``sigb_plan_create``), then one ``sigb_render`` per block.  PyTorch
only owns device buffers and the CUDA stream; every kernel is in libsigb200.so.  No CPU fallback.
"""
from __future__ import annotations

import ctypes
import json
import typing

import numpy as np

from signals_b200 import _lib, plan as plan_mod
from signals_b200 import chain as chain_mod
from signals_b200.chain import (BadShape, BlockLoc, ChainLayerError, FilterDesignError, FilterIndexError,
                                UnsupportedGraph)


class _ShapeError(BadShape):
    def __init__(self, msg):                      # message comes from the library
        ChainLayerError.__init__(self, msg)


def _raise(status: int, where: str):
    detail = _lib.last_error() or _lib.strerror(status)
    if status == _lib.SIGB_ESHAPE:
        raise _ShapeError(detail)
    if status == _lib.SIGB_EINDEX:
        raise FilterIndexError(detail)
    if status == _lib.SIGB_ECRIT:
        raise FilterDesignError(detail)
    if status == _lib.SIGB_EUNSUPPORTED:
        raise UnsupportedGraph(detail)
    if status == _lib.SIGB_EINVAL:
        raise ValueError(f'{where}: {detail}')
    if status == _lib.SIGB_ENOMEM:
        raise MemoryError(f'{where}: {detail}')
    raise RuntimeError(f'{where}: {detail} (status {status}); signals_b200 has no CPU fallback')


def _torch():
    import torch
    return torch


class CompiledPlan:
    """A compiled graph: owns the ``sigb_plan`` handle and the device buffers bound to it."""

    def __init__(self, records: plan_mod.GraphRecords, device=None):
        self.records = records
        self.channels = records.channels
        self.rate = records.rate
        self._lib = _lib.lib()
        self._keep: list = []
        self._windows: dict = {}
        handle = ctypes.c_void_p()
        data = records.data
        dptr = data.ctypes.data_as(ctypes.POINTER(ctypes.c_double)) if data.size else None
        st = self._lib.sigb_plan_create(records.node_array(), len(records.nodes), records.root, dptr, data.size,
                                        records.channels, records.rate, ctypes.byref(handle))
        if st != _lib.SIGB_OK:
            _raise(st, 'sigb_plan_create')
        self.handle = handle
        self.device = device
        self._bound = False

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, 'handle', None):
            self._lib.sigb_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:   # noqa: BLE001 - interpreter shutdown
            pass

    # -- introspection ----------------------------------------------------------------------
    def describe(self) -> dict:
        need = self._lib.sigb_plan_describe(self.handle, None, 0)
        buf = ctypes.create_string_buffer(int(need))
        self._lib.sigb_plan_describe(self.handle, buf, need)
        return json.loads(buf.value.decode())

    def set_option(self, key: str, value: int):
        st = self._lib.sigb_plan_set_option(self.handle, key.encode(), int(value))
        if st != _lib.SIGB_OK:
            _raise(st, 'sigb_plan_set_option')

    @property
    def launch_count(self) -> int:
        return int(self._lib.sigb_plan_launch_count(self.handle))

    def last_kernel_ms(self) -> float:
        ms = ctypes.c_float()
        st = self._lib.sigb_plan_last_kernel_ms(self.handle, ctypes.byref(ms))
        if st != _lib.SIGB_OK:
            _raise(st, 'sigb_plan_last_kernel_ms')
        return float(ms.value)

    def reset(self):
        self._lib.sigb_state_reset(self.handle)

    @property
    def graph_launches(self) -> int:
        """CUDA graphs launched by ``render_block`` so far."""
        return int(self._lib.sigb_plan_graph_launches(self.handle))

    def read_tap(self, index: int, frames: int) -> typing.Optional[np.ndarray]:
        """The block tap ``index`` saw during the most recent request, copied from the buffer the render left in HBM;
        ``None`` when the library cannot serve it (request rendered in several slabs, tap on a Buffer source)."""
        ch = ctypes.c_int32()
        st = self._lib.sigb_plan_read_tap(self.handle, int(index), None, 0, ctypes.byref(ch))
        if st != _lib.SIGB_OK:
            return None
        out = np.empty((frames, ch.value), dtype=np.float32)
        st = self._lib.sigb_plan_read_tap(self.handle, int(index), ctypes.c_void_p(out.ctypes.data), ch.value, None)
        return out if st == _lib.SIGB_OK else None

    # -- rendering --------------------------------------------------------------------------
    def _bind_buffers(self):
        if self._bound:
            return
        torch = _torch()
        for idx, node in self.records.buffers.items():
            s = node.samples
            if not hasattr(s, 'is_cuda'):
                s = torch.from_numpy(np.ascontiguousarray(s, dtype=np.float32))
            t = s.to(device=self.device or 'cuda', dtype=torch.float32).contiguous()
            self._keep.append(t)
            st = self._lib.sigb_plan_bind_buffer(self.handle, idx, ctypes.c_void_p(t.data_ptr()), t.shape[0])
            if st != _lib.SIGB_OK:
                _raise(st, 'sigb_plan_bind_buffer')
        self._bound = True

    def bind_window(self, buffer_node, samples, first_row: int):
        """Streaming source: ``samples`` (CUDA float32 ``(rows, channels)``) holds frames
        ``[first_row, first_row + rows)`` of ``buffer_node``; re-bind before each render of a stream."""
        self._bind_buffers()
        for idx, node in self.records.buffers.items():
            if node is buffer_node:
                assert samples.is_cuda and samples.is_contiguous() and samples.shape[1] == node.channels
                self._windows[idx] = samples          # keeps the memory alive while it is bound
                st = self._lib.sigb_plan_bind_buffer_window(self.handle, idx, ctypes.c_void_p(samples.data_ptr()),
                                                            int(first_row), int(samples.shape[0]))
                if st != _lib.SIGB_OK:
                    _raise(st, 'sigb_plan_bind_buffer_window')
                return
        raise ValueError('bind_window: not a Buffer node of this plan')

    def render_device(self, position: int, frames: int, out=None):
        """Render into a CUDA float32 tensor ``(frames, channels)`` on torch's current stream."""
        torch = _torch()
        if not torch.cuda.is_available():
            raise RuntimeError('signals_b200: no CUDA device -- the block render has no CPU fallback')
        self._bind_buffers()
        if out is None:
            out = torch.empty((frames, self.channels), dtype=torch.float32, device=self.device or 'cuda')
        assert out.is_cuda and out.dtype == torch.float32 and out.shape[0] >= frames and out.stride(1) == 1
        stream = torch.cuda.current_stream(out.device).cuda_stream
        with torch.cuda.device(out.device):
            st = self._lib.sigb_render(self.handle, int(position), int(frames), ctypes.c_void_p(out.data_ptr()),
                                       int(out.stride(0)), ctypes.c_void_p(stream))
        if st != _lib.SIGB_OK:
            _raise(st, 'sigb_render')
        return out

    def render_host(self, position: int, frames: int, out: typing.Optional[np.ndarray] = None):
        """Render into host memory (numpy float32 array or a pinned torch tensor): device->host
        copies are pipelined with the render inside libsigb200 (``sigb_render_host``)."""
        torch = _torch()
        if not torch.cuda.is_available():
            raise RuntimeError('signals_b200: no CUDA device -- the block render has no CPU fallback')
        self._bind_buffers()
        if out is None:
            out = np.empty((frames, self.channels), dtype=np.float32)
        if frames == 0:
            return out                                # an empty block: nothing to render, the stream position stands
        if isinstance(out, np.ndarray):
            assert out.dtype == np.float32 and out.strides[1] == 4
            ptr, ld = out.ctypes.data, out.strides[0] // 4
        else:
            assert out.dtype == torch.float32 and not out.is_cuda and out.stride(1) == 1
            ptr, ld = out.data_ptr(), out.stride(0)
        dev = self.device if self.device is not None else torch.cuda.current_device()
        with torch.cuda.device(dev):
            # work queued on torch's current stream (the upload of a bound Buffer ...) is ordered before the render
            after = torch.cuda.current_stream(dev).cuda_stream
            st = self._lib.sigb_render_host(self.handle, int(position), int(frames), ctypes.c_void_p(ptr), int(ld),
                                            ctypes.c_void_p(after))
        if st != _lib.SIGB_OK:
            _raise(st, 'sigb_render_host')
        return out

    def render_block(self, position: int, frames: int, out: np.ndarray):
        """The audio callback's block (``SinkDevice._callback``, /root/reference/src/signals/chain/dev.py:167-179):
        one captured CUDA graph launch per contiguous block of a given length, delivered into ``out`` (host float32,
        unit channel stride) -- ``sigb_render_block``."""
        if not self._bound:
            self._bind_buffers()
        if frames == 0:
            return out
        assert out.dtype == np.float32 and out.strides[1] == 4
        st = self._lib.sigb_render_block(self.handle, int(position), int(frames), ctypes.c_void_p(out.ctypes.data),
                                         out.strides[0] // 4)
        if st != _lib.SIGB_OK:
            _raise(st, 'sigb_render_block')
        return out


class Engine:
    """Plan cache keyed by (root emitter, channels, rate); recompiles when the graph changes."""

    def __init__(self, device=None, result_dtype=np.float32):
        self.device = device
        self.result_dtype = result_dtype
        self._plans: dict = {}

    def compile(self, emitter, channels: int, rate: int, frames: int = 0) -> CompiledPlan:
        return CompiledPlan(plan_mod.lower(emitter, channels, rate, frames), device=self.device)

    def plan_for(self, emitter, channels: int, rate: int, frames: int = 0) -> CompiledPlan:
        """The cached plan of ``emitter``, recompiled when the graph under it was edited.

        A dirty flag, not a walk: every edit of a ``signals_b200.chain`` graph bumps ``chain.graph_epoch()``
        (port (dis)connected, state attribute assigned, state replaced), so the per-block check is one integer
        compare plus ``np.array_equal`` on the plan's ``Fixed`` arrays (in-place writes into them do not pass through
        a setter).  Graphs holding foreign node objects (the reference's own classes) have no hooks and are
        fingerprinted by ``plan.signature`` as before."""
        key = (id(emitter), int(channels), int(rate))
        hit = self._plans.get(key)
        if hit is not None:
            sig, compiled, _ = hit
            if compiled.records.tracked:
                if sig == chain_mod.graph_epoch() and all(st.value is arr and np.array_equal(arr, snap)
                                                          for st, arr, snap in compiled.records.fixed_values):
                    return compiled
            elif sig == plan_mod.signature(emitter):
                return compiled
            compiled.close()
        epoch = chain_mod.graph_epoch()
        compiled = self.compile(emitter, channels, rate, frames)
        sig = epoch if compiled.records.tracked else plan_mod.signature(emitter)
        self._plans[key] = (sig, compiled, emitter)
        return compiled

    def render(self, emitter, loc: BlockLoc) -> np.ndarray:
        """One block request: returns a host ``(frames, channels)`` array (float32 by default -- the
        device format, /root/reference/src/signals/chain/dev.py:178 casts to it anyway)."""
        frames, channels = loc.shape
        compiled = self.plan_for(emitter, channels, loc.rate, frames)
        if frames == 0:
            return np.zeros((0, channels), dtype=self.result_dtype)
        out = compiled.render_device(loc.position, frames)
        return out.cpu().numpy().astype(self.result_dtype, copy=False)

    def serve_taps(self, emitter, loc: BlockLoc, rendered: typing.Optional[np.ndarray] = None) -> int:
        """Deliver to every enabled tap under ``emitter`` (Wave / Spec / FileWriter) the block it would have seen
        in the reference's recursion.  ``rendered`` = the host block the request at ``loc`` has just produced: taps at
        the root get that block, interior taps get theirs from the buffer the same launch left in HBM
        (``sigb_plan_read_tap``) -- no second render.  Without ``rendered`` (or when the library cannot serve a
        tap) the tap's own input is rendered.  Returns the number of taps served."""
        frames, channels = loc.shape
        compiled = self.plan_for(emitter, channels, loc.rate, frames)
        served = 0
        for tap, creq, index in compiled.records.taps:
            src = tap.inputs_by_port.get('input')
            if src is None or not getattr(tap.get_state(), 'enabled', True):
                continue
            block = None
            if index is None:
                block = rendered                      # a tap at the root sees the rendered block itself
            elif rendered is not None:
                block = compiled.read_tap(index, frames)   # kept aside in HBM by the same launch
            if block is None:                         # no block at hand: a render of the tap's own input
                tap_loc = BlockLoc(position=loc.position, rate=loc.rate, shape=type(loc.shape)(frames=frames, channels=creq))
                block = self.render(src, tap_loc)
            tap.deliver(loc.position, loc.rate, np.broadcast_to(block, (frames, creq)))
            served += 1
        return served

    def render_device(self, emitter, loc: BlockLoc, out=None):
        frames, channels = loc.shape
        return self.plan_for(emitter, channels, loc.rate, frames).render_device(loc.position, frames, out)

    def clear(self):
        for _, compiled, _ in self._plans.values():
            compiled.close()
        self._plans.clear()


_default: typing.Optional[Engine] = None


def default_engine() -> Engine:
    global _default
    if _default is None:
        _default = Engine()
    return _default


def render(emitter, position: int, frames: int, channels: int, rate: int = 48000) -> np.ndarray:
    """Convenience root pull: what ``sink.input.request(BlockLoc(...))`` returns."""
    from signals_b200.chain import Shape
    return default_engine().render(emitter, BlockLoc(position=position, rate=rate,
                                                     shape=Shape(frames=frames, channels=channels)))
