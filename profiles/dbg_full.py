import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from pflotran_b200 import abi, synth, reactive_transport as rt
from oracle.pyoracle import Oracle
from common import workload_cells, rel_err, RTOL
n=200000
w, cells = workload_cells('hanford300a_eq', n)
res={}
for k in (1,3):
    rx = rt.Reaction(w.tables); rz = rt.Realization(rx, n); rz.set_react_kernel(k)
    for f, v in w.base.items(): rz.broadcast(f, v)
    rz.set_cell_scalars(porosity=cells['porosity'], temp=cells['temp'], pres=cells['pres'])
    rz.upload('MNRL_VOLFRAC', cells['volfrac'])
    xg = cells['tran_xx'].copy()
    it, fl = rz.RTReact(xg, 3600.0)
    res[k]=(it,fl,xg, rz.download('TOTAL'))
    print('kernel',k,'flags hist',np.unique(fl,return_counts=True),'iters max',it.max())
it1,fl1,x1,t1=res[1]; it3,fl3,x3,t3=res[3]
d=np.where((it1!=it3)|(fl1!=fl3))[0]
print('cells differing in iters/flags between TPC and lane:', d[:20], len(d))
print('max rel err xx lane vs tpc', rel_err(x3,x1).max(), 'total', rel_err(t3,t1).max())
bad=np.where((fl3!=1)&(fl3!=2))[0]; print('bad cells lane', bad, fl3[bad], it3[bad], 'tpc', fl1[bad], it1[bad])
sample = np.sort(np.random.default_rng(11).choice(n, 3000, replace=False))
sel=np.union1d(sample, np.union1d(d[:50], bad))
sub = {k: (v[sel] if v.ndim == 1 else (v[sel] if k == 'tran_xx' else v[:, sel])) for k, v in cells.items()}
st_o = synth.host_state(w, sub); xo = sub['tran_xx'].copy()
it_o, fl_o = Oracle(w.tables).react(st_o, xo, 3600.0, nthreads=8, maxit=10000)
print('oracle vs lane iters equal', (it_o==it3[sel]).all(), 'flags', (fl_o==fl3[sel]).all(), 'n diff', (it_o!=it3[sel]).sum())
dd=np.where((it_o!=it3[sel])|(fl_o!=fl3[sel]))[0]
print('diff cells', sel[dd], 'oracle', it_o[dd], fl_o[dd], 'lane', it3[sel][dd], fl3[sel][dd], 'tpc', it1[sel][dd], fl1[sel][dd])
print('xx relerr lane vs oracle', rel_err(x3[sel], xo).max(), 'TOTAL', rel_err(t3[:,sel], st_o['TOTAL']).max())
e=rel_err(t3[:,sel], st_o['TOTAL']); print('TOTAL worst', np.unravel_index(e.argmax(), e.shape), e.max())
