"""Dry-run harness for bench.py in the CPU suite -- TEST INFRASTRUCTURE ONLY.

`python -m cusim.fake_cuda <bench.py arguments>` runs bench.py's own `main()` with (a) torch's CUDA entry points replaced by
CPU stand-ins ("device" tensors are CPU tensors, events are wall-clock stamps, the nccl backend becomes gloo) and (b) the
package's C-ABI binding pointed at the emulated library (tests/cusim).  Nothing it prints is a measurement; the point is
that every line of the bench's control flow (workload construction, timed loops, the end-to-end legs, the JSON line, the
collectives of the N > 1 path) executes before the one run on the B200 that counts.  The NVLink peer-memory exchange cannot
be emulated across processes (its handles are raw pointers), so N > 1 dry-runs use --exchange nccl."""
import os
import runpy
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.dirname(HERE)):
    if p not in sys.path:
        sys.path.insert(0, p)


def install():
    import torch
    import torch.distributed as dist

    class Event:
        def __init__(self, enable_timing=False):
            self.t = 0.0

        def record(self, stream=None):
            self.t = time.perf_counter()

        def elapsed_time(self, other):
            return 1e3 * (other.t - self.t)

        def synchronize(self):
            pass

    class Stream:
        cuda_stream = 0

    torch.cuda.is_available = lambda: True
    torch.cuda.set_device = lambda d: None
    torch.cuda.current_device = lambda: 0
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.current_stream = lambda *a, **k: Stream()
    torch.cuda.device_count = lambda: 8
    torch.cuda.Event = Event
    torch.Tensor.cuda = lambda self, *a, **k: self.clone()      # a device copy never aliases the host array

    def drop_device(fn):
        def wrapped(*a, **k):
            if str(k.get("device", "")).startswith("cuda"):
                k.pop("device")
            return fn(*a, **k)
        return wrapped
    for name in ("empty", "tensor", "zeros", "full", "empty_like"):
        setattr(torch, name, drop_device(getattr(torch, name)))
    real_init = dist.init_process_group

    def init_process_group(backend=None, *a, **k):
        k.pop("device_id", None)
        return real_init("gloo", *a, **k)
    dist.init_process_group = init_process_group

    import cusim
    import ndb200
    ndb200._cabi._lib = cusim.lib()
    # CPU tensors stand in for device tensors: route them to the device entry points like real CUDA tensors
    from networkdynamics_jl_b200 import network
    real_addr = network._addr

    def addr(x):
        a, dev, n = real_addr(x)
        if hasattr(x, "data_ptr"):
            dev = True
        return a, dev, n
    network._addr = addr


if __name__ == "__main__":
    install()
    sys.argv = [os.path.join(ROOT, "bench.py")] + sys.argv[1:]
    runpy.run_path(os.path.join(ROOT, "bench.py"), run_name="__main__")
