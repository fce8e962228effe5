"""TEST HARNESS (CPU): run the Python host mirror and the `-m gpu` test functions WITHOUT a GPU.

`install()` (called from tests/conftest.py when MJB_TEST_EMU=1) does two things in the test process only:
  * points `mjmpc_b200._lib` at tests/hostcheck/libmjmpc_b200_emu.so -- the product's own .cu sources (kernels and
    launch code) built for the host by gen_lib_emu.py, same extern "C" entry points, "device" pointers = host
    pointers;
  * makes torch hand out CPU tensors wherever the code asks for a CUDA device, and replaces the stream / event
    calls by synchronous stand-ins.  CUDA graphs cannot be emulated: controllers report themselves un-graphable
    and run the same step eagerly.
Nothing under mjmpc_b200/ refers to this module: the product path still refuses to run without CUDA and its
extension (tests/test_host_cpu.py checks that).  What this buys: every line of host glue, argument marshalling,
launch code and kernel source written between GPU sessions is executed before it reaches the GPU box."""
import contextlib
import ctypes as C
import importlib.util
import os
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
_installed = False


def build_lib():
    import sys
    hc = os.path.join(ROOT, "tests", "hostcheck")
    if hc not in sys.path:
        sys.path.insert(0, hc)
    spec = importlib.util.spec_from_file_location("gen_lib_emu", os.path.join(hc, "gen_lib_emu.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    return gen.build()


class _Stream:
    cuda_stream = 0

    def __init__(self, *a, **k):
        pass

    def wait_stream(self, other):
        pass

    def wait_event(self, ev):
        pass

    def synchronize(self):
        pass

    def record_event(self, ev=None):
        ev = ev or _Event()
        ev.record()
        return ev


class _Event:
    def __init__(self, *a, **k):
        self.t = None

    def record(self, stream=None):
        self.t = time.perf_counter()

    def synchronize(self):
        pass

    def wait(self, stream=None):
        pass

    def query(self):
        return True

    def elapsed_time(self, other):
        return (other.t - self.t) * 1e3


def install():
    global _installed
    if _installed:
        return
    _installed = True
    import torch
    from mjmpc_b200 import _lib
    from mjmpc_b200.control.controller import Controller

    L = C.CDLL(build_lib())
    L.mjb_last_error.restype = C.c_char_p
    for name in _lib.EXPORTS:
        getattr(L, name)
    _lib._setup_restypes(L)
    _lib._lib = L

    cpu = torch.device("cpu")

    def dev(d):
        if isinstance(d, str):
            return "cpu" if d.startswith("cuda") else d
        if isinstance(d, int):
            return cpu
        if isinstance(d, torch.device) and d.type == "cuda":
            return cpu
        return d

    def factory(fn):
        def wrapped(*a, **k):
            if "device" in k:
                k["device"] = dev(k["device"])
            k.pop("pin_memory", None)
            return fn(*a, **k)
        return wrapped

    for name in ("empty", "zeros", "ones", "full", "tensor", "as_tensor", "eye", "randn", "rand", "arange",
                 "empty_like", "zeros_like", "ones_like", "full_like"):
        setattr(torch, name, factory(getattr(torch, name)))

    real_to = torch.Tensor.to

    def to(self, *a, **k):
        a = tuple(dev(x) if isinstance(x, (str, torch.device)) else x for x in a)
        if "device" in k:
            k["device"] = dev(k["device"])
        return real_to(self, *a, **k)

    torch.Tensor.to = to
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.Tensor.pin_memory = lambda self, *a, **k: self

    stream = _Stream()
    torch.cuda.is_available = lambda: True
    torch.cuda.device_count = lambda: 1
    torch.cuda.current_device = lambda: 0
    torch.cuda.set_device = lambda d: None
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.current_stream = lambda *a, **k: stream
    torch.cuda.Stream = _Stream
    torch.cuda.Event = _Event
    torch.cuda.stream = lambda s: contextlib.nullcontext()
    torch.cuda.empty_cache = lambda: None
    # CUDA graphs: not emulated -- enable_cuda_graph returns False and the step runs eagerly (same kernels)
    Controller._graphable = lambda self: False


def active():
    return _installed
