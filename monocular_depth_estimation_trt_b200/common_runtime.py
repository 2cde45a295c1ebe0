"""Host side of one inference: pinned/device buffer pairs, H2D -> enqueue -> D2H -> sync, stage timer.

Drop-in for the reference's core/common_runtime.py (same public names, argument meaning and
error behaviour) with TensorRT taken out: `engine` / `context` are the objects of
`monocular_depth_estimation_trt_b200.engine`, which expose the slice of
tensorrt.ICudaEngine / IExecutionContext this module touches.

  reference (core/common_runtime.py)            here
  ------------------------------------------    -----------------------------------------
  check_cuda_err / cuda_call        :41-56      same contract: RuntimeError on a CUDA error
  HostDeviceMem                     :59-108     same: .host (pinned numpy view, checked setter),
                                                .device (int), .nbytes, .free()
  allocate_buffers                  :131-175    same signature and return tuple
  free_buffers                      :179-182    same
  StageTimer                        :196-238    same: mark(i, stream) / read() / last / free()
  do_inference                      :268-275    same: returns [flat pinned numpy views]

Memory and streams come from the CUDA runtime through cuda-python, as in the reference, so the
`stream` handed around is a cudart stream and `device` attributes are plain integer addresses.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional, Sequence, Union

import numpy as np

try:                                    # cuda-python >= 12.8 layout
    from cuda.bindings import driver as cuda, runtime as cudart
except ImportError:                     # pragma: no cover - older cuda-python
    from cuda import cuda, cudart       # type: ignore

_H2D = cudart.cudaMemcpyKind.cudaMemcpyHostToDevice
_D2H = cudart.cudaMemcpyKind.cudaMemcpyDeviceToHost


def check_cuda_err(err) -> None:
    """Raise RuntimeError for anything but success; unknown status types are errors too."""
    if isinstance(err, cuda.CUresult):
        if err != cuda.CUresult.CUDA_SUCCESS:
            raise RuntimeError(f"Cuda Error: {err}")
        return
    if isinstance(err, cudart.cudaError_t):
        if err != cudart.cudaError_t.cudaSuccess:
            raise RuntimeError(f"Cuda Runtime Error: {err}")
        return
    raise RuntimeError(f"Unknown error type: {err}")


def cuda_call(call):
    """Unwrap cuda-python's (status, *results) tuples."""
    status, results = call[0], call[1:]
    check_cuda_err(status)
    return results[0] if len(results) == 1 else results


def _volume(shape: Sequence[int]) -> int:
    n = 1
    for s in shape:
        n *= int(s)
    return n


class HostDeviceMem:
    """One binding: `size` elements of `dtype` in pinned host memory and the same bytes on the device."""

    def __init__(self, size: int, dtype: Optional[np.dtype] = None):
        dtype = np.dtype(dtype) if dtype is not None else np.dtype(np.uint8)
        self._nbytes = int(size) * dtype.itemsize
        self._host_ptr = cuda_call(cudart.cudaMallocHost(self._nbytes))
        raw = (ctypes.c_uint8 * self._nbytes).from_address(int(self._host_ptr))
        self._host = np.frombuffer(raw, dtype=np.uint8).view(dtype)
        self._device = int(cuda_call(cudart.cudaMalloc(self._nbytes)))

    @property
    def host(self) -> np.ndarray:
        return self._host

    @host.setter
    def host(self, data: Union[np.ndarray, bytes]) -> None:
        if isinstance(data, np.ndarray):
            if data.size > self._host.size:
                raise ValueError(
                    f"Tried to fit an array of size {data.size} into host memory of size {self._host.size}")
            np.copyto(self._host[:data.size], data.reshape(-1), casting="safe")
            return
        if self._host.dtype != np.uint8:
            raise TypeError("raw bytes can only be assigned to a uint8 binding")
        payload = np.frombuffer(data, dtype=np.uint8)
        if payload.size > self._nbytes:
            raise ValueError(f"Tried to fit {payload.size} bytes into host memory of {self._nbytes} bytes")
        self._host[:payload.size] = payload

    @property
    def device(self) -> int:
        return self._device

    @property
    def nbytes(self) -> int:
        return self._nbytes

    def __repr__(self) -> str:
        return f"HostDeviceMem(nbytes={self._nbytes}, dtype={self._host.dtype}, device=0x{self._device:x})"

    __str__ = __repr__

    def free(self) -> None:
        cuda_call(cudart.cudaFree(self._device))
        cuda_call(cudart.cudaFreeHost(self._host_ptr))
        self._host = np.empty(0, dtype=self._host.dtype)


def _shape_override(name: str, shape: Sequence[int], output_shape):
    """`output_shape` is a {binding: shape} dict, or one shape used only where the engine's own
    shape is unusable (dynamic, or a volume <= 1)."""
    if output_shape is None:
        return None
    if isinstance(output_shape, dict):
        return output_shape.get(name)
    usable = all(int(s) >= 0 for s in shape) and _volume(shape) > 1
    return None if usable else output_shape


def allocate_buffers(engine, output_shape=None, profile_idx: Optional[int] = None):
    """-> (inputs, outputs, bindings, stream), one HostDeviceMem per I/O tensor in engine order."""
    inputs: List[HostDeviceMem] = []
    outputs: List[HostDeviceMem] = []
    bindings: List[int] = []
    stream = cuda_call(cudart.cudaStreamCreate())
    for i in range(engine.num_io_tensors):
        name = engine.get_tensor_name(i)
        shape = (engine.get_tensor_shape(name) if profile_idx is None
                 else engine.get_tensor_profile_shape(name, profile_idx)[-1])
        if profile_idx is None and any(int(s) < 0 for s in shape):
            raise ValueError(f"Binding {name} has dynamic shape, but no profile was specified.")
        size = _volume(shape)
        override = _shape_override(name, shape, output_shape)
        if override is not None:
            size = _volume(override)
        mem = HostDeviceMem(size, np.dtype(engine.get_tensor_dtype(name)))
        bindings.append(int(mem.device))
        (inputs if engine.get_tensor_mode(name) == engine.TensorIOMode.INPUT else outputs).append(mem)
    return inputs, outputs, bindings, stream


def free_buffers(inputs: List[HostDeviceMem], outputs: List[HostDeviceMem], stream) -> None:
    for mem in list(inputs) + list(outputs):
        mem.free()
    cuda_call(cudart.cudaStreamDestroy(stream))


def memcpy_host_to_device(device_ptr: int, host_arr: np.ndarray) -> None:
    cuda_call(cudart.cudaMemcpy(device_ptr, host_arr.ctypes.data, host_arr.size * host_arr.itemsize, _H2D))


def memcpy_device_to_host(host_arr: np.ndarray, device_ptr: int) -> None:
    cuda_call(cudart.cudaMemcpy(host_arr.ctypes.data, device_ptr, host_arr.size * host_arr.itemsize, _D2H))


class StageTimer:
    """Four CUDA events on the inference stream: h2d | compute | d2h, in milliseconds.

    GPU-side durations only; they do not add up to the wall clock core.bench reports, which also
    holds launch overhead and the final synchronise.  Events are created once and reused."""

    STAGES = ("h2d_ms", "compute_ms", "d2h_ms")

    def __init__(self):
        self._events = [cuda_call(cudart.cudaEventCreate()) for _ in range(len(self.STAGES) + 1)]
        self.last: Dict[str, float] = {}

    def mark(self, i: int, stream) -> None:
        cuda_call(cudart.cudaEventRecord(self._events[i], stream))

    def read(self) -> Dict[str, float]:
        """Only valid once the stream has synchronised."""
        self.last = {
            stage: float(cuda_call(cudart.cudaEventElapsedTime(self._events[k], self._events[k + 1])))
            for k, stage in enumerate(self.STAGES)
        }
        return self.last

    def free(self) -> None:
        for ev in self._events:
            cuda_call(cudart.cudaEventDestroy(ev))
        self._events = []


def _do_inference_base(inputs, outputs, stream, execute_async_func, timer: Optional[StageTimer] = None):
    def mark(i):
        if timer is not None:
            timer.mark(i, stream)

    mark(0)
    for inp in inputs:
        cuda_call(cudart.cudaMemcpyAsync(inp.device, inp.host.ctypes.data, inp.nbytes, _H2D, stream))
    mark(1)
    execute_async_func()
    mark(2)
    for out in outputs:
        cuda_call(cudart.cudaMemcpyAsync(out.host.ctypes.data, out.device, out.nbytes, _D2H, stream))
    mark(3)
    cuda_call(cudart.cudaStreamSynchronize(stream))
    if timer is not None:
        timer.read()
    return [out.host for out in outputs]


def do_inference(context, engine, bindings, inputs, outputs, stream, timer: Optional[StageTimer] = None):
    """One forward, blocking: returns flat views of the outputs' pinned memory (overwritten by the next call)."""
    for i in range(engine.num_io_tensors):
        context.set_tensor_address(engine.get_tensor_name(i), bindings[i])
    return _do_inference_base(inputs, outputs, stream,
                              lambda: context.execute_async_v3(stream_handle=stream), timer)
