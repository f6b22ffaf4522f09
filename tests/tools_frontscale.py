import sys, ctypes, time, os
sys.path.insert(0, os.environ.get('GRAFT_REPO_ROOT','/root/repo'))
import numpy as np
from lepton_b200.codec import lib, _Buffer
import bench
L=lib()
L.lepb200_host_frontend_seconds.restype=ctypes.c_double
L.lepb200_host_frontend_seconds.argtypes=[ctypes.POINTER(_Buffer), ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int32)]
distinct=bench.make_corpus(16)
n=int(sys.argv[1]) if len(sys.argv)>1 else 512
bufs=(_Buffer*n)(); keep=[]
for i in range(n):
    b=np.frombuffer(distinct[i%16],dtype=np.uint8); keep.append(b); bufs[i].data=b.ctypes.data; bufs[i].len=len(b)
tot=sum(len(distinct[i%16]) for i in range(n))
for th in [1,8,32,64,128]:
    if th>os.cpu_count(): break
    m=min(n, max(16, th*4))
    err=ctypes.c_int32()
    dt=L.lepb200_host_frontend_seconds(bufs,m,th,ctypes.byref(err))
    print('threads',th,'images',m,'sec %.3f'%dt,'ms/img/thread %.1f'%(dt*th/m*1e3),'MB/s %.0f'%(tot/n*m/dt/1e6),'err',err.value, flush=True)
