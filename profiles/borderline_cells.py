import os, sys
import numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from pflotran_b200 import abi, synth, reactive_transport as rt
from oracle.pyoracle import Oracle
for name, n in (('hanford300a_stoich', 300000), ('hanford300a_kinsrf', 300000)):
    w = synth.Workload(name)
    cells = synth.make_cells(w, 0, n, seed=88003)
    st_o = synth.host_state(w, cells); st_g = st_o.copy()
    rx = rt.Reaction(w.tables); rz = rt.Realization(rx, n); rz.upload_host_state(st_g)
    xg = cells['tran_xx'].copy(); it_g, fl_g = rz.RTReact(xg, 3600.0, abi.RXN_DT_CONSISTENT)
    xo = cells['tran_xx'].copy(); it_o, fl_o = Oracle(w.tables).react(st_o, xo, 3600.0, abi.RXN_DT_CONSISTENT, maxit=10000, nthreads=16)
    bad = np.where((it_g != it_o) | (fl_g != fl_o))[0]
    for c in bad:
        print(name, 'cell', c, 'gpu its/flags', it_g[c], hex(fl_g[c]), 'oracle', it_o[c], hex(fl_o[c]), 'max rel diff of result', float(np.max(np.abs(xg[c]-xo[c])/np.abs(xo[c]))))
