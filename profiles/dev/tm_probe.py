import numpy as np, sys
sys.path.insert(0,'/root/repo')
from vcf2prot_b200 import GpuEngine
eng=GpuEngine(0)
ref=np.frombuffer(bytes(range(65,91))*400,dtype=np.uint8).copy()
eng.set_reference(ref,"tensormap")
z=lambda *a: np.asarray(a,dtype=np.uint64)
import os
dst=int(os.environ.get("DST","0"))
tasks=np.asarray([(5,1024,dst,0)],np.uint32)
out,ms=eng.execute_batch(z(0,1),tasks,None,np.zeros(0,np.uint8),z(0,0),z(0,dst+1024))
print("ok", bytes(out[dst:dst+20]), bytes(ref[5:25]))
