"""Kernel time and HBM-roofline fraction of the flux-side kernels (k_flux_residual, k_flux_jacobian,
k_flux_coefs) on a structured nx*ny*nz block, outputs resident on the device.
usage: python profiles/bench_flux.py [workload] [nx ny nz]"""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from pflotran_b200 import synth, reactive_transport as rt
from flux_common import structured_connections

name = sys.argv[1] if len(sys.argv) > 1 else 'hanford300a_eq'
nx, ny, nz = (int(a) for a in sys.argv[2:5]) if len(sys.argv) > 4 else (128, 128, 64)
n_cells = nx * ny * nz
w = synth.Workload(name)
n = w.tables.naqcomp
cells = synth.make_cells(w, 0, n_cells)
rx = rt.Reaction(w.tables)
rz = rt.Realization(rx, n_cells)
for f, v in w.base.items():
    rz.broadcast(f, v)
rz.set_cell_scalars(porosity=cells['porosity'], temp=cells['temp'], pres=cells['pres'])
if w.tables.nkinmnrl:
    rz.upload('MNRL_VOLFRAC', cells['volfrac'])
rz.materialize('DTOTAL')
rng = np.random.default_rng(7)
xx = np.ascontiguousarray(w.base['PRI_MOLAL'][None, :] * np.exp(0.1 * rng.standard_normal((n_cells, w.ncomp))))
rz.RTUpdateAuxVars(xx, True)
conn, nghosted, nlocal, active = structured_connections(nx, ny, nz, n)
nconn = len(conn['id_up'])
cs = rt.ConnectionSet(rz, conn['id_up'], conn['id_dn'], nlocal)
nnzb = cs.nnz_blocks
d_r = rz.device_alloc(nlocal * n * 8)
d_v = rz.device_alloc(nnzb * n * n * 8)
peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))).get('hbm_gbs', 6535.7) if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else 6535.7
out = {'workload': name, 'grid': [nx, ny, nz], 'cells': n_cells, 'connections': nconn, 'jacobian_blocks': nnzb, 'hbm_peak_gbs': peak}


def timed(fn, reps=6):
    ms = []
    for _ in range(reps):
        fn()
        ms.append(rz.last_kernel_ms())
    return float(np.median(ms[2:]))


structure_bytes = nnzb * 8 + (nlocal + 1) * 4 + nlocal * 4      # col + ent, row_ptr, l2g
legs = {
    'flux_coefs': (lambda: cs.TFluxCoef(conn['area'], conn['velocity'], conn['disp']), nconn * (n + 2) * 8 + 2 * n * nconn * 8),
    'flux_residual': (lambda: rz.RTResidualFlux_device(cs, d_r), n * n_cells * 8 + 2 * n * nconn * 8 + structure_bytes + nlocal * n * 8),
    'flux_jacobian': (lambda: rz.RTJacobianFlux_device(cs, d_v), n * n * n_cells * 8 + 2 * n * nconn * 8 + structure_bytes + nnzb * n * n * 8),
}
for label, (fn, nbytes) in legs.items():
    ms = timed(fn)
    out[label] = {'kernel_ms': ms, 'algorithmic_bytes': nbytes, 'gbs': nbytes / (ms * 1e-3) / 1e9, 'hbm_frac': nbytes / (ms * 1e-3) / 1e9 / peak,
                  'rows_per_s': nlocal / (ms * 1e-3)}
chk = np.zeros(1024)
rz.device_copy(chk, d_v, chk.nbytes, 1)
out['checksum'] = float(np.abs(chk).sum())
print(json.dumps(out))
