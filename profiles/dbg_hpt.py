import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from pflotran_b200 import abi, synth, reactive_transport as rt
from oracle.pyoracle import Oracle
from common import workload_cells, rel_err, RTOL
name='hpt_calcite'; n=3000
w, cells = workload_cells(name, n)
st_o = synth.host_state(w, cells); st_g = st_o.copy()
rx = rt.Reaction(w.tables); rz = rt.Realization(rx, n); rz.upload_host_state(st_g)
orc = Oracle(w.tables)
rng = np.random.default_rng(7)
xx = np.ascontiguousarray(w.base['PRI_MOLAL'][None, :] * np.exp(0.1 * rng.standard_normal((n, w.ncomp))))
orc.update_auxvars(st_o, xx, True, nthreads=8); rz.RTUpdateAuxVars(xx, True)
a_o = orc.fixed_accum(st_o, xx, nthreads=8); a_g = rz.RTUpdateFixedAccumulation(xx)
r_o, j_o = orc.residual_jacobian(st_o, 1800.0, nthreads=8)
r_g, j_g = rz.RTResidualJacobianNonFlux(1800.0)
rs = np.maximum(np.abs(r_o), 1e-12 * np.abs(r_o).max(axis=1, keepdims=True))
e = np.abs(r_g - r_o) / np.maximum(rs, 1e-300)
k = np.unravel_index(e.argmax(), e.shape)
print('max err', e.max(), 'at', k, 'r_o', r_o[k[0]], 'r_g', r_g[k[0]])
rz.download_host_state(st_g)
print('total', st_o['TOTAL'][:, k[0]], 'accum', a_o[k[0]]/1800.0, 'rate', st_o['MNRL_RATE'][:, k[0]], st_g['MNRL_RATE'][:, k[0]])
print('n cells over tol', (e.max(axis=1) > RTOL).sum())
