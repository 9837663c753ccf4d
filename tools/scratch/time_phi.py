import sys, numpy as np, torch, ctypes as C
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from helpers import load_golden
import lightspinner_b200.engine as E
from lightspinner_b200 import _capi
p, _ = load_golden('c2_falc_cah')
nc = 256
e = E.MaliEngine(p, nc, max_upload_chunk=nc)
e.upload_device_phi([p] * nc)
N = e.mt.Nspace
aD = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(p['aDamp'], (nc,) + p['aDamp'].shape))).cuda()
vB = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(p['vBroad'], (nc,) + p['vBroad'].shape))).cuda()
vL = torch.zeros((nc, 1, N), dtype=torch.float64).cuda() + 1500.0
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for rep in range(3):
    ev[0].record()
    _capi.check(e.lib.mali_compute_phi(e._handle, C.byref(e.bufs), 0, nc, C.c_void_p(aD.data_ptr()), C.c_void_p(vB.data_ptr()), C.c_void_p(vL.data_ptr()), e._stream()))
    ev[1].record(); torch.cuda.synchronize()
    print('compute_phi for %d C2 columns: %.3f ms' % (nc, ev[0].elapsed_time(ev[1])))
