# needs the debug build: SGMC_TC_DEBUG=1 python -c "import __graft_entry__ as g; g.build()"
import sys, ctypes as C, numpy as np
sys.path.insert(0,'/root/repo')
from jax_sgmc_b200 import device, ops, _lib
from jax_sgmc_b200.device import DeviceArray as DA
device.set_device(0)
import os
Cc,d,n,N=[int(v) for v in os.environ.get("SHAPE","4096,1024,1024,100000").split(",")]
X,y,_=ops.synth_logistic_data(0,N,d)
theta=DA.from_numpy((np.random.default_rng(0).standard_normal((Cc,d))*0.3).astype(np.float32))
idx=DA((n,),np.int32); dk=[DA.from_numpy(ops.prng_key(0)),DA((2,),np.uint32)]
ops.minibatch_draw(dk[0],dk[1],idx,N)
U,var,g=DA((Cc,),np.float32),DA((Cc,),np.float32),DA((Cc,d),np.float32)
spec=ops.glm_spec("logistic",d,0,prior="gaussian",prior_off=0,prior_size=d,prior_scale=10.0)
GR=int(os.environ.get("GRAD","1")); lib=_lib.load(); lib.sgmc_debug_tc_timers.argtypes=[C.c_void_p]
for path in ("tc_parity","tc_throughput"):
  ws=ops.glm_workspace(Cc,n,d,path)
  for rep in range(3):
    ops.glm_potential_grad(spec,theta,X,y,idx,N,U,var,(g if GR else None),workspace=ws,path=path)
    device.synchronize()
    t=(C.c_ulonglong*10)(); lib.sgmc_debug_tc_timers(t)
    a=[int(t[i])-int(t[0]) for i in range(8)]
  print(path,("GEMM2" if GR else "GEMM1")+" ns from start: setup",a[1],"mainloop_done",a[2],"tmem_ld",a[4],"staged",a[5],"chunk0 done",a[6],"loops done",a[7],"end",a[3])
