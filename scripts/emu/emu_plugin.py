"""DEVELOPMENT TOOL (see include/cuda_runtime.h): pytest plugin of the dry run.  Presents the host model to the tests
as "cuda": torch factory calls / .cuda() / .to("cuda") land on the CPU (whose pointers the model's cudaMemcpy
understands), torch.cuda.* bookkeeping calls are no-ops, and ctypes.CDLL("libcudart.so") is the model's runtime."""
import contextlib
import ctypes
import os

import torch
from torch.overrides import TorchFunctionMode

assert os.environ.get("CHIMP_EMU") == "1", "the emulation plugin is only for scripts/emu/run.sh"
ROOT = os.getcwd()
EMU_LIB = os.path.join(ROOT, "badchimp-cpp_b200", "libchimp_b200.so")


def _is_cuda(d):
    if isinstance(d, torch.device):
        return d.type == "cuda"
    if isinstance(d, str):
        return d.startswith("cuda")
    return False


class _ToCpu(TorchFunctionMode):
    def __torch_function__(self, func, types, args=(), kwargs=None):
        kwargs = dict(kwargs or {})
        if _is_cuda(kwargs.get("device")):
            kwargs["device"] = "cpu"
        kwargs.pop("pin_memory", None)
        name = getattr(func, "__name__", "")
        if name == "cuda":
            return args[0]
        if name == "to":
            args = tuple("cpu" if _is_cuda(a) else a for a in args)
        if name in ("pin_memory",):
            return args[0]
        if name == "is_pinned":
            return True
        return func(*args, **kwargs)


_CDLL = ctypes.CDLL
_mode = _ToCpu()
_mode.__enter__()


_rt = None


def _runtime():
    global _rt
    if _rt is None:
        _rt = _CDLL(EMU_LIB)
        _rt.cudaStreamSynchronize.argtypes = [ctypes.c_void_p]
    return _rt


class _Stream:
    def __init__(self, *a, **k):
        self.cuda_stream = a[0] if a else 0

    def synchronize(self):
        if self.cuda_stream:
            _runtime().cudaStreamSynchronize(self.cuda_stream)

    def wait_stream(self, s):
        pass


class _Event:
    def __init__(self, *a, **k):
        import time
        self._t = 0.0
        self._time = time

    def record(self, stream=None):
        self._t = self._time.perf_counter()

    def synchronize(self):
        pass

    def elapsed_time(self, other):
        return (other._t - self._t) * 1e3


torch.cuda.is_available = lambda: True
torch.cuda.device_count = lambda: int(os.environ.get("LOCAL_WORLD_SIZE", "1"))
torch.cuda.current_device = lambda: 0
torch.cuda.set_device = lambda d: None
torch.cuda.synchronize = lambda *a, **k: _runtime().cudaDeviceSynchronize()
torch.cuda.empty_cache = lambda: None
torch.cuda.can_device_access_peer = lambda a, b: True


@contextlib.contextmanager
def _on_stream(s):
    # torch work "on" an engine stream runs eagerly here: it is in stream order once the stream has drained
    if s is not None:
        s.synchronize()
    yield


torch.cuda.stream = _on_stream
torch.cuda.ExternalStream = _Stream
torch.cuda.Stream = _Stream
torch.cuda.Event = _Event
torch.cuda.current_stream = lambda *a, **k: _Stream(0)
torch.cuda.get_device_name = lambda *a: "host model of the engine (scripts/emu)"
torch.cuda.mem_get_info = lambda *a: (64 << 30, 64 << 30)

class _PatchedCDLL(_CDLL):
    def __init__(self, name, *a, **k):
        if isinstance(name, str) and os.path.basename(name).startswith("libcudart.so"):
            name = EMU_LIB
        super().__init__(name, *a, **k)


ctypes.CDLL = _PatchedCDLL


# N ranks as N processes (torchrun): the collectives run over gloo on CPU tensors; a collective "on a stream" first lets
# every stream of this rank's model drain, which is the order NCCL would have had on that stream
import torch.distributed as _dist  # noqa: E402

_init_pg = _dist.init_process_group


def _init_process_group(backend=None, *a, **k):
    k.pop("device_id", None)
    return _init_pg("gloo", *a, **k)


_dist.init_process_group = _init_process_group


def _drained(fn):
    def wrapper(*a, **k):
        _runtime().cudaDeviceSynchronize()
        return fn(*a, **k)
    wrapper.__name__ = getattr(fn, "__name__", "collective")
    return wrapper


# (isend / irecv themselves stay untouched: P2POp compares them by identity)
for _name in ("all_reduce", "all_gather", "all_gather_object", "gather_object", "barrier", "batch_isend_irecv", "broadcast"):
    if hasattr(_dist, _name):
        setattr(_dist, _name, _drained(getattr(_dist, _name)))
