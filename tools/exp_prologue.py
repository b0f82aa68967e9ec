import sys; sys.path.insert(0,'/root/repo')
import numpy as np
from jax_sgmc_b200 import device, ops
from jax_sgmc_b200.device import DeviceArray as DA, Event, Stream
device.set_device(0); s=Stream.create(); device.set_current_stream(s)
C,P,R=4096,1024,6
outs=[DA.zeros((C,P)) for _ in range(R)]
kk=ops.prng_keys(range(C))
for mode in (0,99):
  ops.set_option(1,mode)
  for i in range(R): ops.normal_like(kk,[P],out=outs[i])
  s.sync(); e0,e1=Event(),Event(); e0.record(s)
  for i in range(20*R): ops.normal_like(kk,[P],out=outs[i%R])
  e1.record(s); e1.sync()
  print("mode",mode,"normal_like us",e0.elapsed_ms(e1)*1e3/(20*R))
