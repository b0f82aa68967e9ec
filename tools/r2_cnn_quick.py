import sys, json; sys.path.insert(0, '/root/repo')
import bench
from jax_sgmc_b200 import _lib, device
from jax_sgmc_b200.device import Stream
class A: pass
_lib.load(); device.set_device(0); s = Stream.create(); device.set_current_stream(s)
ctl = bench.Control(0, 1)
print(json.dumps(bench.bench_cnn(A(), ctl, None, s), indent=1))
