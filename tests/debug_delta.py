import numpy as np, torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import ops
import gpu_util as G
N,T,F,theta=26,40,50,9
rng=np.random.default_rng(N+T+F+theta)
x=rng.normal(size=(N,T,F)).astype('float32')
want=ops.delta_fwd(x,theta)
ldx,ldy=52,152
xp=np.zeros((N*T,ldx),'float32'); xp[:,:F]=x.reshape(N*T,F)
dx,dy=G.dev(xp),G.zeros((N*T,ldy))
G.call('ipavsr_delta_fwd',dx.data_ptr(),ldx,dy.data_ptr(),ldy,N,T,F,theta,1,G.stream())
got=G.host(dy)[:,:3*F].reshape(N,T,3*F)
mm=np.argwhere(got!=want)
print('mismatches',len(mm),'of',got.size)
print('by block', [(got[:,:,k*F:(k+1)*F]!=want[:,:,k*F:(k+1)*F]).sum() for k in range(3)])
for (n,t,c) in mm[:12]:
    a,b=got[n,t,c],want[n,t,c]
    print(n,t,c,a,b,(a.view(np.int32)-b.view(np.int32)))
# host emulation of the kernel formula with long double
def emu(A):
    tt=np.arange(T); d=np.zeros_like(A)
    for th in range(1,theta+1):
        diff=(A[:,np.minimum(tt+th,T-1)]-A[:,np.maximum(tt-th,0)]).astype(np.float32)
        r=np.float64(1.0)/np.float64(2.0*th)
        s=(d.astype(np.longdouble)+diff.astype(np.longdouble)*np.longdouble(r)).astype(np.float64)
        d=s.astype(np.float32)
    return d
d_e=emu(x); a_e=emu(d_e)
print('emu vs oracle d mism', (d_e!=want[:,:,F:2*F]).sum(), 'emu vs gpu d mism', (d_e!=got[:,:,F:2*F]).sum())
# which theta step diverges for first mismatch in d block
md=np.argwhere(got[:,:,F:2*F]!=want[:,:,F:2*F])
if len(md):
    n,t,f=md[0]
    acc=np.float32(0)
    for th in range(1,theta+1):
        hi=x[n,min(t+th,T-1),f]; lo=x[n,max(t-th,0),f]; diff=np.float32(hi-lo)
        term=np.float64(th)*np.float64(diff)/np.float64(2*th*th)
        s_ref=np.float64(acc)+term
        s_fma=np.float64(np.longdouble(acc)+np.longdouble(diff)*np.longdouble(np.float64(1)/np.float64(2*th)))
        print(th, repr(diff), repr(term), repr(s_ref), repr(s_fma), repr(np.float32(s_ref)), repr(np.float32(s_fma)))
        acc=np.float32(s_ref)
    print('gpu', repr(got[n,t,F+f]), 'want', repr(want[n,t,F+f]))
