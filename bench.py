#!/usr/bin/env python
"""bench.py — reaction cell-updates/s of the batched RReact path on N B200s (one rank per GPU).

  python bench.py --gpus 1 --steps 5 --warmup 3            # our arm (CUDA, through the C ABI)
  python bench.py --impl reference --steps 3 --warmup 1    # reference arm: the CPU path on host cores

A step = one RTReact pass (reference src/pflotran/reactive_transport.F90:1605-1806) over this
rank's batch of synthetic cells (pflotran_b200/synth.py).  Every step starts from the
same inputs (transported totals in tran_xx, initial guess = base state), restored on the device
inside the timed region, so no step is cheaper than the first.

Timed regions (CUDA events on the library's stream, barrier + synchronize on both sides, max
over ranks):
  value  : K x (device-side input restore + react kernel), inputs resident in HBM
  e2e    : K x (pinned host tran_xx -> H2D -> restore + kernel -> D2H of free-ion result, iteration
           counts and flags) through the public host-buffer call Realization.RTReact (rxn_react_batch moves the batch
           in chunks so that the PCIe copies overlap the kernel)
  roofline.achieved : the react kernel alone (rxn_last_kernel_ms, CUDA events around the launch)
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from pflotran_b200 import abi, synth  # noqa: E402

METRIC = 'reaction cell-updates/sec'
UNIT = 'cell-updates/s'
DEFAULT_CELLS = {'hanford300a_eq': 10_000_000, 'hanford300a_mr': 5_000_000, 'calcite': 1_000_000,
                 'hpt_calcite': 1_000_000, 'ascem': 1_000_000, 'scco2_brine': 1_000_000}
WORKLOAD_DESC = {
    'hanford300a_eq': 'hanford/300A U(VI) chemistry (15 primaries, 88 complexes, 2 kinetic minerals, equilibrium surface '
                      'complexation; deck regression_tests/default/543/543_hanford_srfcplx_base.in), synthetic grid',
    'hanford300a_mr': '300A chemistry with 50-rate multirate surface complexation (543_hanford_srfcplx_mr.in)',
    'calcite': 'example_problems/100_100_100 calcite chemistry (4 primaries, 5 complexes, 1 kinetic mineral)',
    'hpt_calcite': 'geothermal-hpt.dat calcite chemistry, per-cell T,P dependent logK',
    'scco2_brine': 'aqueous chemistry of the MPHASE CO2 deck regression_tests/default/scco2/mphase/mphase_chem.in (8 primaries with CO2(aq) in the '
                   'basis, 12 complexes, kinetic quartz + calcite, 1 molal NaCl brine), without the supercritical phase',
    'ascem': 'example_problems/ascem_chemistry (BASELINE config 1: 22 primaries, 164 complexes, calcite kinetics; database savannah_river.dat)',
}
# DRAM bytes per cell-update of the react kernel: a CONSTANT taken from the committed `ncu --set full` capture of the named kernel
# (dram__bytes_read.sum + dram__bytes_write.sum over the cells of the profiled launch), not measured in the run; reported only
# when the kernel the run used is the captured one (same name prefix), else null
NCU_DRAM_BYTES_PER_CELL = {
    'hanford300a_eq': [('tensor-memory N=15 cells/CTA=128 warps/cell=3', (905.745152e6 + 1396.868e6) / 600000, 'profiles/r02_g_tm_g3.metrics.txt'),
                       ('tensor-memory N=15 cells/CTA=128 warps/cell=4', (932.234496e6 + 1454.456e6) / 600000, 'profiles/r02_f_tm_g4.metrics.txt'),
                       ('resident-lane N=15 cells/CTA=64 lanes/cell=2', (758.277120e6 + 1237.866e6) / 600000,
                        'profiles/r01_r9_lane_g2.metrics.csv')],
}


def ncu_traffic(workload, kernel_info, ncells):
    for prefix, per_cell, src in NCU_DRAM_BYTES_PER_CELL.get(workload, []):
        if kernel_info.startswith(prefix):
            return per_cell * ncells, 'constant from the ncu --set full capture of this kernel (%s), per cell x cells of one launch; not measured in this run' % src
    return None, 'no ncu capture of this kernel/workload committed'


RESET_FIELDS = ['PRI_MOLAL', 'PRI_ACT_COEF', 'SEC_MOLAL', 'SEC_ACT_COEF', 'LN_ACT_H2O', 'TOTAL_SORB_EQ', 'FREE_SITE_CONC',
                'EQIONX_REF_CATION_SORBED_CONC']


# ------------------------------------------------------------------------------------------------
def work_model(t, iters_sum: float, ncells: int):
    """Algorithmic work per SURVEY.md 8d: flop-equivalents (W = 20 per transcendental) and
    compulsory bytes for `ncells` cell-updates whose Newton iteration counts sum to iters_sum."""
    naq, ncomp, ncplx, nkin = t.naqcomp, t.ncomp, t.neqcplx, t.nkinmnrl
    S = int(sum(t.eqcplxspecid[k, 0] for k in range(ncplx)))
    S2 = int(sum(int(t.eqcplxspecid[k, 0]) ** 2 for k in range(ncplx)))
    nsrf, nrxn = t.nsrfcplx, t.nsrfcplxrxn
    nrate = t.kinmr_max_nrate if t.nkinmrsrfcplxrxn else 0
    srf_n2 = int(sum(int(t.srfcplxspecid[k, 0]) ** 2 for k in range(nsrf)))
    F_rtotal = 7 * S + 2 * ncplx + 2 * S2 + naq + naq * naq
    F_lu = (2.0 / 3.0) * ncomp ** 3 + 4 * ncomp ** 2
    F_sorb = 12 * srf_n2 + 4 * naq * (1 + (1 if nrate else 0)) if nsrf else 0
    kin_n2 = int(sum(int(t.kinmnrlspecid[k, 0]) ** 2 for k in range(nkin)))
    F_min = 20 * nkin + 2 * kin_n2
    k = iters_sum / ncells
    a = k if t.act_coef_update_frequency == 2 else (1 if t.act_coef_update_frequency == 1 else 0)
    F = (k + 1) * (F_rtotal + F_sorb) + k * (F_lu + 4 * naq * naq + F_min) + a * 15 * (naq + ncplx)
    its_site = 2
    Nt = (k + 1) * (2 * naq + ncplx + S + nsrf * its_site) + k * (nkin * 2 + 2 * naq) + a * (naq + ncplx + 1)
    flop_eq = F + 20.0 * Nt
    rd = ncomp + naq + ncplx + 7 + 2 * nkin + nrxn + naq * nrate * t.nkinmrsrfcplxrxn
    wr = ncomp + naq + naq + ncplx + naq + ncplx + nkin + nrxn + nsrf + naq + naq * t.nkinmrsrfcplxrxn + 1
    return {'flop_eq_per_cell': flop_eq, 'flops_per_cell': F, 'transcendentals_per_cell': Nt,
            'bytes_per_cell': 8.0 * (rd + wr), 'mean_newton_iterations': k}


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), 'measured (MEASURED_PEAKS.json)'
    return {'hbm_gbs': 6650.0}, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q, '--format=csv,noheader,nounits',
                                          '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ts, line in self.lines:
            f = [x.strip() for x in line.split(',')]
            if len(f) < 7 or not (t0 <= ts <= t1 + 0.2):
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ------------------------------------------------------------------------------------------------
def cpu_reference_rate(w, ncells_sample: int, start: int, dt: float, threads: int, repeats: int = 1):
    """Times the CPU restatement of the reference's RTReact loop (oracle/, kind "port": no Fortran
    compiler exists in this image) on `ncells_sample` cells with `threads` host threads."""
    from oracle.pyoracle import Oracle
    cells = synth.make_cells(w, start, ncells_sample)
    orc = Oracle(w.tables)
    best = None
    iters = None
    for _ in range(repeats):
        st = synth.host_state(w, cells)
        xx = cells['tran_xx'].copy()
        t0 = time.perf_counter()
        iters, flags = orc.react(st, xx, dt, abi.RXN_DT_CONSISTENT, maxit=10000, nthreads=threads)
        el = time.perf_counter() - t0
        best = el if best is None else min(best, el)
    return ncells_sample / best, best, float(iters.mean())


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    w = synth.Workload(args.workload)
    threads = os.cpu_count() or 1
    # size the per-step sample for ~6 s of CPU work
    rate0, _, _ = cpu_reference_rate(w, 4096 * max(1, threads // 4), 0, args.dt, threads)
    sample = int(max(4096, min(DEFAULT_CELLS[args.workload], rate0 * 6.0)) // 4096 * 4096)
    for _ in range(args.warmup):
        cpu_reference_rate(w, min(sample, 8192 * threads), 0, args.dt, threads)
    t_tot = 0.0
    k_mean = 0.0
    for s in range(args.steps):
        rate, el, km = cpu_reference_rate(w, sample, 0, args.dt, threads)
        t_tot += el
        k_mean = km
    value = sample * args.steps / t_tot
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * t_tot / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': WORKLOAD_DESC[args.workload], 'name': args.workload, 'cells_per_step': sample,
                   'dt_s': args.dt, 'dt_mode': 'DT_CONSISTENT', 'mean_newton_iterations': k_mean},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                         'sample': 'first %d cells of the synthetic workload per step; C++ restatement of the reference '
                                   'RTReact loop (oracle/rxn_oracle.cpp), %d host threads, static partition' % (sample, threads)},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# Global-implicit mode (BASELINE config 4): one step = the per-Newton-iteration cell loops of the global-implicit transport
# solve: RTUpdateAuxVars with activity-coefficient update (reactive_transport.F90:3790-3846, 3620-3700) on a new free-ion
# iterate, then the accumulation + reaction residual and Jacobian blocks of RTResidualNonFlux / RTJacobianNonFlux
# (:2545-2586, 2735-2758, 3342-3389, 3445-3465) for every cell.
GI_METRIC = 'global-implicit reaction residual+Jacobian blocks/sec'
GI_UNIT = 'cell-blocks/s'
GI_DEFAULT_CELLS = {'hpt_calcite': 4_000_000, 'calcite': 4_000_000, 'hanford300a_eq': 1_000_000, 'hanford300a_mr': 500_000, 'ascem': 500_000,
                    'scco2_brine': 2_000_000}


def gi_work_model(t):
    """Algorithmic work of one global-implicit step per cell (same accounting as work_model: SURVEY.md 8d, W = 20
    flop-equivalents per transcendental): RTotal + RTotalSorb are evaluated twice (RTUpdateAuxVars, and again inside the block
    evaluation, which keeps no naq^2 array in HBM), activity coefficients once, minerals once."""
    naq, ncplx, nkin = t.naqcomp, t.neqcplx, t.nkinmnrl
    S = int(sum(t.eqcplxspecid[k, 0] for k in range(ncplx)))
    S2 = int(sum(int(t.eqcplxspecid[k, 0]) ** 2 for k in range(ncplx)))
    nsrf, nrxn = t.nsrfcplx, t.nsrfcplxrxn
    nrate = t.kinmr_max_nrate if t.nkinmrsrfcplxrxn else 0
    srf_n2 = int(sum(int(t.srfcplxspecid[k, 0]) ** 2 for k in range(nsrf)))
    kin_n2 = int(sum(int(t.kinmnrlspecid[k, 0]) ** 2 for k in range(nkin)))
    F_rtotal = 7 * S + 2 * ncplx + 2 * S2 + naq + naq * naq
    F_sorb = 12 * srf_n2 + 4 * naq * (1 + (1 if nrate else 0)) if nsrf else 0
    F_min = 20 * nkin + 2 * kin_n2
    F = 2 * (F_rtotal + F_sorb) + F_min + 4 * naq * naq + 15 * (naq + ncplx) + 2 * naq * nrate * t.nkinmrsrfcplxrxn
    Nt = 2 * (2 * naq + ncplx + S + nsrf * 2) + 2 * nkin + (naq + ncplx + 1)
    if t.logK_mode != 0:                     # per-cell logK: 17-term / 5-term fit per reaction, evaluated in both kernels
        nrx = ncplx + nkin + nsrf
        F += 2 * 40 * nrx
        Nt += 2 * 3
    mr = naq * nrate * t.nkinmrsrfcplxrxn
    rd1 = naq + ncplx + 7 + nkin + nrxn                                    # xx, lagged sec_molal, scalars, volfrac, free sites
    wr1 = 3 * naq + 2 * ncplx + naq + nrxn + nsrf + 1                      # pri_molal, total, gamma; sec_molal, gamma_k; sorbed; ...
    rd2 = 2 * naq + ncplx + 7 + 2 * nkin + nrxn + mr
    wr2 = naq + naq * naq + naq + ncplx + nrxn + nkin + naq + nsrf + naq * t.nkinmrsrfcplxrxn
    return {'flop_eq_per_cell': F + 20.0 * Nt, 'flops_per_cell': F, 'transcendentals_per_cell': Nt,
            'bytes_per_cell': 8.0 * (rd1 + wr1 + rd2 + wr2)}


def gi_inputs(w, cells):
    """Free-ion iterate of the step: the base state's molalities under the workload's per-cell perturbation."""
    return np.ascontiguousarray(w.base['PRI_MOLAL'][None, :] * (cells['tran_xx'] / w.base['TOTAL'][None, :]))


def cpu_gi_rate(w, ncells_sample, dt, threads):
    from oracle.pyoracle import Oracle
    cells = synth.make_cells(w, 0, ncells_sample)
    orc = Oracle(w.tables)
    st = synth.host_state(w, cells)
    xx = gi_inputs(w, cells)
    t0 = time.perf_counter()
    orc.update_auxvars(st, xx, True, nthreads=threads)
    orc.residual_jacobian(st, dt, nthreads=threads)
    el = time.perf_counter() - t0
    return ncells_sample / el, el


class GiBench:
    """State + buffers of one global-implicit benchmark instance (used by --mode gi and by the extra configs of the default run)."""

    def __init__(self, rt, name, n, device, start=0):
        self.rt, self.n = rt, n
        self.w = w = synth.Workload(name)
        self.t = t = w.tables
        self.cells = cells = synth.make_cells(w, start, n)
        self.rx = rt.Reaction(t, device=device)
        self.rz = rz = rt.Realization(self.rx, n)
        for f, v in w.base.items():
            rz.broadcast(f, v)
        rz.set_cell_scalars(porosity=cells['porosity'], temp=cells['temp'], pres=cells['pres'])
        if t.nkinmnrl:
            rz.upload('MNRL_VOLFRAC', cells['volfrac'])
        nc = t.ncomp
        self.xx_host = rt.pinned_empty((n, nc))
        self.xx_host[:] = gi_inputs(w, cells)
        self.res_host = rt.pinned_empty((n, nc))
        self.jac_host = rt.pinned_empty((n, nc * nc))
        self.d_xx = rz.device_alloc(n * nc * 8)
        self.d_res = rz.device_alloc(n * nc * 8)
        self.d_jac = rz.device_alloc(n * nc * nc * 8)
        rz.device_copy(self.d_xx, self.xx_host, n * nc * 8, 0)
        self.h2d = n * nc * 8
        self.d2h = n * (nc + nc * nc) * 8

    def restore(self):
        for f in RESET_FIELDS:
            if self.rx.field_rows(f):
                self.rz.broadcast(f, self.w.base[f])

    def step_device(self, dt):
        self.restore()
        self.rz.RTUpdateAuxVars_device(self.d_xx, True)
        k1 = self.rz.last_kernel_ms()
        self.rz.RTResidualJacobianNonFlux_device(self.n, dt, self.d_res, self.d_jac)
        return k1 + self.rz.last_kernel_ms()

    def step_e2e(self, dt):
        self.restore()
        self.rz.RTUpdateAuxVars(self.xx_host, True)
        self.rz.RTResidualJacobianNonFlux(dt, res=self.res_host, jac=self.jac_host)

    def measure(self, steps, warmup, dt, barrier=lambda: None):
        rz = self.rz
        for _ in range(max(warmup, 3)):
            self.step_device(dt)
        barrier()
        rz.timer_start()
        kern = [self.step_device(dt) for _ in range(steps)]
        dev_ms = rz.timer_stop()
        barrier()
        self.step_e2e(dt)
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            self.step_e2e(dt)
        barrier()
        e2e_s = time.perf_counter() - t0
        return dev_ms, statistics.mean(kern), e2e_s

    def rooflines(self, kern_ms, fp64_peak, hbm_peak):
        wm = gi_work_model(self.t)
        ks = kern_ms * 1e-3
        f_ach = wm['flop_eq_per_cell'] * self.n / ks / 1e12
        b_ach = wm['bytes_per_cell'] * self.n / ks / 1e9
        fp = {'bound': 'fp64', 'achieved': f_ach, 'peak': fp64_peak, 'unit': 'TFLOP/s', 'frac': f_ach / fp64_peak, 'traffic': None,
              'kernel_ms': kern_ms, 'flop_eq_per_cell': wm['flop_eq_per_cell'],
              'kernels': 'update_auxvars + residual/Jacobian blocks (sum of the two launches, CUDA events)'}
        hb = {'bound': 'hbm', 'achieved': b_ach, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': b_ach / hbm_peak, 'traffic': None,
              'kernel_ms': kern_ms, 'bytes_per_cell': wm['bytes_per_cell']}
        # the binding limit is the one that would take longer at its peak
        return (fp, hb) if f_ach / fp64_peak >= b_ach / hbm_peak else (hb, fp)


def run_gi(args):
    import torch
    import torch.distributed as dist
    from pflotran_b200 import reactive_transport as rt
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the product path has no CPU fallback')
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        tt = torch.tensor([x], dtype=torch.float64, device='cuda')
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    name = args.workload if args.workload_given else 'hpt_calcite'
    n = args.cells or GI_DEFAULT_CELLS[name]
    g = GiBench(rt, name, n, local_rank, start=rank * n)
    fp64_peak = g.rz.probe_fp64_tflops()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    l0 = rt.launch_count()
    tw0 = time.perf_counter()
    dev_ms, kern_ms, e2e_s = g.measure(args.steps, args.warmup, args.gi_dt, barrier)
    tw1 = time.perf_counter()
    launches = rt.launch_count() - l0
    clocks = sampler.stop(tw0, tw1)
    dev_ms = max_over_ranks(dev_ms); kern_ms = max_over_ranks(kern_ms); e2e_s = max_over_ranks(e2e_s)
    if rank == 0:
        peaks, peak_src = measured_peaks()
        r1, r2 = g.rooflines(kern_ms, fp64_peak, peaks['hbm_gbs'])
        total = n * world
        line = {
            'metric': GI_METRIC, 'value': total * args.steps / (dev_ms * 1e-3), 'unit': GI_UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': dev_ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': WORKLOAD_DESC[name] + '; global-implicit step: RTUpdateAuxVars (activity update) + reaction '
                                   'residual and Jacobian block of every cell', 'name': name, 'mode': 'gi', 'cells_per_gpu': n,
                       'total_cells': total, 'dt_s': args.gi_dt,
                       'l2_policy': 'cell state + output blocks per GPU larger than L2 (%.1f GB); no flush needed'
                                    % (n * gi_work_model(g.t)['bytes_per_cell'] / 1e9)},
            'e2e': {'value': total * (args.steps) / e2e_s, 'unit': GI_UNIT, 'h2d_bytes_per_step': g.h2d, 'd2h_bytes_per_step': g.d2h},
            'gpu_launches': int(launches * world), 'clocks': clocks,
            'roofline': dict(r1, peak_source=peak_src if r1['bound'] == 'hbm' else 'DFMA probe measured in this run (rxn_probe_fp64)'),
            'roofline_other': r2,
        }
        threads = os.cpu_count() or 1
        rate0, _ = cpu_gi_rate(g.w, 4096 * max(1, threads // 4), args.gi_dt, threads)
        sample = int(max(4096, min(n, rate0 * 10.0)) // 4096 * 4096)
        rate, el = cpu_gi_rate(g.w, sample, args.gi_dt, threads)
        line['cpu_baseline'] = {'value': rate, 'unit': GI_UNIT, 'cores': threads, 'kind': 'port',
                                'sample': 'first %d cells of the same workload, %.1f s, C++ restatement of RTUpdateAuxVars + the '
                                          'accumulation/reaction loops of RTResidual/RTJacobian (oracle/)' % (sample, el)}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_gi_reference(args):
    if int(os.environ.get('RANK', '0')) != 0:
        return
    name = args.workload if args.workload_given else 'hpt_calcite'
    w = synth.Workload(name)
    threads = os.cpu_count() or 1
    rate0, _ = cpu_gi_rate(w, 4096 * max(1, threads // 4), args.gi_dt, threads)
    sample = int(max(4096, min(GI_DEFAULT_CELLS[name], rate0 * 6.0)) // 4096 * 4096)
    for _ in range(args.warmup):
        cpu_gi_rate(w, min(sample, 8192 * threads), args.gi_dt, threads)
    t_tot = 0.0
    for _ in range(args.steps):
        _, el = cpu_gi_rate(w, sample, args.gi_dt, threads)
        t_tot += el
    value = sample * args.steps / t_tot
    print(json.dumps({
        'impl': 'reference', 'metric': GI_METRIC, 'value': value, 'unit': GI_UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * t_tot / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': WORKLOAD_DESC[name], 'name': name, 'mode': 'gi', 'cells_per_step': sample, 'dt_s': args.gi_dt},
        'cpu_baseline': {'value': value, 'unit': GI_UNIT, 'cores': threads, 'kind': 'port',
                         'sample': 'first %d cells per step, C++ restatement (oracle/), %d host threads' % (sample, threads)},
        'e2e': {'value': value, 'unit': GI_UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0}))


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from pflotran_b200 import reactive_transport as rt

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the product path has no CPU fallback')
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    w = synth.Workload(args.workload)
    t = w.tables
    n = args.cells or DEFAULT_CELLS[args.workload]
    if args.scaling == 'strong':           # the total stays that of one GPU: every rank takes its contiguous share
        n = (n + world - 1) // world
    ncomp = t.ncomp
    start = rank * n                       # every rank owns n cells, globally numbered (weak: n fixed; strong: n = total / ranks)
    cells = synth.make_cells(w, start, n)
    rx = rt.Reaction(t, device=local_rank)
    rz = rt.Realization(rx, n)
    if args.kernel:
        rz.set_react_kernel(args.kernel)
    for f, v in w.base.items():
        rz.broadcast(f, v)
    rz.set_cell_scalars(porosity=cells['porosity'], temp=cells['temp'], pres=cells['pres'])
    if t.nkinmnrl:
        rz.upload('MNRL_VOLFRAC', cells['volfrac'])
    mr_field = w.base.get('KINMR_TOTAL_SORB')

    xx_host = rt.pinned_empty((n, ncomp))
    xx_host[:] = cells['tran_xx']
    xx0_host = cells['tran_xx']
    it_host = rt.pinned_empty((n,), np.int32)
    fl_host = rt.pinned_empty((n,), np.int32)
    nb = n * ncomp * 8
    d_xx0 = rz.device_alloc(nb)
    d_xx = rz.device_alloc(nb)
    d_it = rz.device_alloc(n * 4)
    d_fl = rz.device_alloc(n * 4)
    rz.device_copy(d_xx0, xx0_host, nb, 0)

    def restore():
        for f in RESET_FIELDS:
            if rx.field_rows(f):
                rz.broadcast(f, w.base[f])

    def step_device():
        restore()
        rz.device_copy(d_xx, d_xx0, nb, 2)
        rz.RTReact_device(d_xx, n, args.dt, abi.RXN_DT_CONSISTENT, 0, d_it, d_fl)
        return rz.last_kernel_ms()

    # e2e: RTReact overwrites the caller's (pinned) Vec in place, so every timed step gets its own pre-filled buffer; when
    # that would pin more than a quarter of the node's free memory the one buffer is re-filled inside the timed region instead (and counted)
    e2e_bufs = [xx_host]
    try:
        import psutil
        pin_budget = 0.25 * psutil.virtual_memory().available   # pinned host memory of all ranks of the node together
    except Exception:
        pin_budget = 16 * 2 ** 30
    if nb * (args.steps + 1) * world <= pin_budget:
        for _ in range(args.steps):
            b = rt.pinned_empty((n, ncomp))
            b[:] = xx0_host
            e2e_bufs.append(b)

    def step_e2e(k=0):
        restore()
        if len(e2e_bufs) > 1:
            buf = e2e_bufs[k]
        else:
            buf = xx_host
            buf[:] = xx0_host          # the caller's Vec content for this step (host memcpy, part of the step)
        rz.RTReact(buf, args.dt, abi.RXN_DT_CONSISTENT, iters=it_host, flags=fl_host)

    fp64_peak = rz.probe_fp64_tflops()
    for _ in range(max(args.warmup, 3)):
        step_device()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    l0 = rt.launch_count()
    barrier()
    tw0 = time.perf_counter()
    rz.timer_start()
    kern_ms = []
    for _ in range(args.steps):
        kern_ms.append(step_device())
    dev_ms = rz.timer_stop()
    barrier()
    tw1 = time.perf_counter()
    launches = rt.launch_count() - l0
    clocks = sampler.stop(tw0, tw1)
    dev_ms = max_over_ranks(dev_ms)
    rz.device_copy(it_host, d_it, n * 4, 1)
    rz.device_copy(fl_host, d_fl, n * 4, 1)
    iters_sum = float(it_host.sum(dtype=np.int64))
    bad = int(((fl_host != abi.RXN_EXIT_RESIDUAL) & (fl_host != abi.RXN_EXIT_REL_CHANGE)).sum())
    iters_sum_all = sum_over_ranks(iters_sum)
    bad_all = int(sum_over_ranks(float(bad)))
    kern_ms_max = max_over_ranks(statistics.mean(kern_ms))

    # the same steps with the work order switched off (RXN_NO_REACT_ORDER: the lanes take the batch as it comes), reported beside
    # the default so that the share of the order in `value` is visible in the line
    unordered = None
    if not os.environ.get('RXN_NO_REACT_ORDER') and work_order_active(rz.react_kernel_info(), n, local_rank):
        os.environ['RXN_NO_REACT_ORDER'] = '1'
        step_device()
        barrier()
        rz.timer_start()
        km_u = [step_device() for _ in range(args.steps)]
        dev_ms_u = max_over_ranks(rz.timer_stop())
        del os.environ['RXN_NO_REACT_ORDER']
        unordered = (dev_ms_u, max_over_ranks(statistics.mean(km_u)))

    # end-to-end through the host-buffer call
    xx_host[:] = xx0_host
    step_e2e(0)
    barrier()
    te0 = time.perf_counter()
    for k in range(args.steps):
        step_e2e(k + 1 if len(e2e_bufs) > 1 else 0)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - te0)

    if rank == 0:
        total_cells = n * world
        value = total_cells * args.steps / (dev_ms * 1e-3)
        wm = work_model(t, iters_sum_all, total_cells)
        peaks, peak_src = measured_peaks()
        kern_s = kern_ms_max * 1e-3
        fp64_ach = wm['flop_eq_per_cell'] * n / kern_s / 1e12
        hbm_ach = wm['bytes_per_cell'] * n / kern_s / 1e9
        kinfo = rz.react_kernel_info()
        traffic, traffic_src = ncu_traffic(args.workload, kinfo, n)
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': dev_ms / args.steps, 'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None,
            'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': WORKLOAD_DESC[args.workload], 'name': args.workload, 'cells_per_gpu': n,
                       'total_cells': total_cells, 'dt_s': args.dt, 'dt_mode': 'DT_CONSISTENT',
                       'mean_newton_iterations': wm['mean_newton_iterations'], 'cells_with_nonreference_flags': bad_all,
                       'work_order': ('lanes take the cells sorted by the Newton iteration counts of the previous call (sort inside the timed region); every '
                                      'step of this benchmark repeats the same inputs, so that prediction is exact here; roofline.unordered = the same '
                                      'steps without it') if unordered is not None else 'off (index order) at this batch size / chemistry',
                       'l2_policy': 'inputs (%.1f GB of cell state per GPU) larger than L2; no flush needed'
                                    % (n * wm['bytes_per_cell'] / 1e9),
                       'kernel': kinfo},
            'e2e': {'value': total_cells * args.steps / e2e_s, 'unit': UNIT, 'h2d_bytes_per_step': nb,
                    'd2h_bytes_per_step': nb + 2 * n * 4},
            'gpu_launches': int(launches * world),
            'clocks': clocks,
            'roofline': {'bound': 'fp64', 'achieved': fp64_ach, 'peak': fp64_peak, 'unit': 'TFLOP/s', 'frac': fp64_ach / fp64_peak,
                         'traffic': traffic, 'traffic_source': traffic_src,
                         'note': 'FP64 CUDA-core path: achieved = flop-equivalents (SURVEY.md 8d, W=20 per exp/log/sqrt/pow) x cells '
                                 '/ react-kernel time (CUDA events); peak = DFMA probe measured in this run (rxn_probe_fp64)',
                         'kernel_ms': kern_ms_max, 'flop_eq_per_cell': wm['flop_eq_per_cell'],
                         'transcendentals_per_cell': wm['transcendentals_per_cell'],
                         'unordered': None if unordered is None else {
                             'value': total_cells * args.steps / (unordered[0] * 1e-3), 'kernel_ms': unordered[1],
                             'frac': wm['flop_eq_per_cell'] * n / (unordered[1] * 1e-3) / 1e12 / fp64_peak,
                             'note': 'the same steps with RXN_NO_REACT_ORDER=1 (lanes take the batch in index order)'},
                         # the probe's number beside the arithmetic one: SMs x 64 FP64 lanes x 2 flop x the maximum SM clock
                         # (MEASURED_PEAKS.json has no FP64 entry; NVIDIA's nominal 40 TFLOP/s is this product at 2.1 GHz)
                         'peak_theoretical': theoretical_fp64_tflops(local_rank, clocks),
                         'frac_of_theoretical': (fp64_ach / theoretical_fp64_tflops(local_rank, clocks))
                         if theoretical_fp64_tflops(local_rank, clocks) else None},
            'roofline_hbm': {'bound': 'hbm', 'achieved': hbm_ach, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                             'frac': hbm_ach / peaks['hbm_gbs'],
                             'traffic': traffic, 'traffic_source': traffic_src, 'peak_source': peak_src,
                             'bytes_per_cell': wm['bytes_per_cell']},
        }
        # CPU baseline: bounded sample on the host cores of this box
        threads = os.cpu_count() or 1
        rate0, _, _ = cpu_reference_rate(w, 4096 * max(1, threads // 4), 0, args.dt, threads)
        sample = int(max(4096, min(n, rate0 * 10.0)) // 4096 * 4096)
        rate, el, km = cpu_reference_rate(w, sample, 0, args.dt, threads)
        line['cpu_baseline'] = {'value': rate, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                                'sample': 'first %d cells of the same workload, %.1f s, C++ restatement of the reference loop '
                                          '(no Fortran compiler in the image)' % (sample, el)}
        if world == 1 and not args.no_extra and not args.workload_given and not args.cells:
            line['other_configs'] = extra_configs(rt, rz, local_rank, args, fp64_peak, peaks['hbm_gbs'])
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def work_order_active(kernel_info, n, device):
    """Whether the library orders a single launch of n cells by the previous call's iteration counts (rxn_b200.cu: react_ordered):
    tail-bound chemistries always, chemistries with >= 8 primaries below 16 generations of resident cells; read off the kernel line."""
    import re
    if 'work order:' in kernel_info:
        return True
    m = re.search(r'cells/CTA=(\d+)', kernel_info)
    if 'work order below 16 generations' in kernel_info and m:
        import torch
        return n < 16 * torch.cuda.get_device_properties(device).multi_processor_count * int(m.group(1))
    return False


def theoretical_fp64_tflops(device, clocks):
    """SM count x 64 FP64 FMA lanes per SM x 2 flop x maximum SM clock (None when the clock could not be sampled)."""
    try:
        import torch
        sms = torch.cuda.get_device_properties(device).multi_processor_count
        mhz = clocks.get('sm_max_mhz')
        return sms * 64 * 2 * mhz * 1e6 / 1e12 if mhz else None
    except Exception:
        return None


def extra_configs(rt, rz_main, device, args, fp64_peak, hbm_peak):
    """The other single-GPU BASELINE configurations, measured in the same run after the headline (short: 3 steps each), so
    that the driver's default invocation records them: config 2 (calcite chemistry on the 100^3 grid, operator-split RTReact),
    the chemistry of config 1 (ascem, 22 primaries / 164 complexes) and config 4 (geothermal-hpt chemistry, global-implicit residual/Jacobian blocks).  Same timing rules as the headline."""
    out = []

    def react_config(label, name, n, K, cpu_cells):
        w = synth.Workload(name)
        t = w.tables
        cells = synth.make_cells(w, 0, n)
        rx = rt.Reaction(t, device=device)
        rz = rt.Realization(rx, n)
        for f, v in w.base.items():
            rz.broadcast(f, v)
        rz.set_cell_scalars(porosity=cells['porosity'], temp=cells['temp'], pres=cells['pres'])
        if t.nkinmnrl:
            rz.upload('MNRL_VOLFRAC', cells['volfrac'])
        nb = n * t.ncomp * 8
        it_host = rt.pinned_empty((n,), np.int32); fl_host = rt.pinned_empty((n,), np.int32)
        d_xx0 = rz.device_alloc(nb); d_xx = rz.device_alloc(nb); d_it = rz.device_alloc(n * 4); d_fl = rz.device_alloc(n * 4)
        rz.device_copy(d_xx0, cells['tran_xx'], nb, 0)

        def restore():
            for f in RESET_FIELDS:
                if rx.field_rows(f):
                    rz.broadcast(f, w.base[f])

        def step():
            restore()
            rz.device_copy(d_xx, d_xx0, nb, 2)
            rz.RTReact_device(d_xx, n, args.dt, abi.RXN_DT_CONSISTENT, 0, d_it, d_fl)
            return rz.last_kernel_ms()
        for _ in range(3):
            step()
        rz.timer_start()
        km = [step() for _ in range(K)]
        dev_ms = rz.timer_stop()
        rz.device_copy(it_host, d_it, n * 4, 1)
        bufs = []                                               # RTReact overwrites the caller's Vec: one pre-filled pinned buffer per step
        for _ in range(K + 1):
            b = rt.pinned_empty((n, t.ncomp))
            b[:] = cells['tran_xx']
            bufs.append(b)
        restore()
        rz.RTReact(bufs[K], args.dt, abi.RXN_DT_CONSISTENT, iters=it_host, flags=fl_host)
        t0 = time.perf_counter()
        for k in range(K):
            restore()
            rz.RTReact(bufs[k], args.dt, abi.RXN_DT_CONSISTENT, iters=it_host, flags=fl_host)
        e2e_s = time.perf_counter() - t0
        wm = work_model(t, float(it_host.sum(dtype=np.int64)), n)
        ks = statistics.mean(km) * 1e-3
        fa = wm['flop_eq_per_cell'] * n / ks / 1e12
        threads = os.cpu_count() or 1
        rate, el, _ = cpu_reference_rate(w, cpu_cells * max(1, threads // 4), 0, args.dt, threads)
        return {'config': label, 'metric': METRIC, 'unit': UNIT, 'name': name, 'workload': WORKLOAD_DESC[name],
                'cells': n, 'steps': K, 'value': n * K / (dev_ms * 1e-3),
                'e2e': {'value': n * K / e2e_s, 'unit': UNIT, 'h2d_bytes_per_step': nb, 'd2h_bytes_per_step': nb + 8 * n},
                'roofline': {'bound': 'fp64', 'achieved': fa, 'peak': fp64_peak, 'unit': 'TFLOP/s', 'frac': fa / fp64_peak,
                             'kernel_ms': statistics.mean(km), 'traffic': None},
                'roofline_hbm_frac': wm['bytes_per_cell'] * n / ks / 1e9 / hbm_peak,
                'mean_newton_iterations': wm['mean_newton_iterations'], 'kernel': rz.react_kernel_info(),
                'cpu_baseline': {'value': rate, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'sample': '%.1f s' % el}}

    # config 2 (1M cells, as BASELINE names it) and the chemistry of config 1 (the reference's CPU-runnable case; 22 primaries on
    # the N = 24 resident-lane shape, 500 000 cells: a launch ends with the tail of its damped redox cells)
    for label, name, n, K, cpu_cells in (('BASELINE config 2', 'calcite', DEFAULT_CELLS['calcite'], 10, 65536),
                                         ('BASELINE config 1 chemistry', 'ascem', 500_000, 3, 8192)):
        try:
            out.append(react_config(label, name, n, K, cpu_cells))
        except Exception as e:                                  # the headline line must not be lost to an extra configuration
            out.append({'config': label, 'error': repr(e)})
    try:
        name, n = 'hpt_calcite', GI_DEFAULT_CELLS['hpt_calcite']
        g = GiBench(rt, name, n, device)
        K = 5
        dev_ms, kern_ms, e2e_s = g.measure(K, 3, args.gi_dt)
        r1, r2 = g.rooflines(kern_ms, fp64_peak, hbm_peak)
        threads = os.cpu_count() or 1
        rate, el = cpu_gi_rate(g.w, 65536 * max(1, threads // 4), args.gi_dt, threads)
        out.append({'config': 'BASELINE config 4', 'metric': GI_METRIC, 'unit': GI_UNIT, 'name': name, 'workload': WORKLOAD_DESC[name],
                    'cells': n, 'steps': K, 'value': n * K / (dev_ms * 1e-3),
                    'e2e': {'value': n * K / e2e_s, 'unit': GI_UNIT, 'h2d_bytes_per_step': g.h2d, 'd2h_bytes_per_step': g.d2h},
                    'roofline': r1, 'roofline_other': r2,
                    'cpu_baseline': {'value': rate, 'unit': GI_UNIT, 'cores': threads, 'kind': 'port', 'sample': '%.1f s' % el}})
    except Exception as e:
        out.append({'config': 'BASELINE config 4', 'error': repr(e)})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='hanford300a_eq', choices=sorted(DEFAULT_CELLS))
    ap.add_argument('--cells', type=int, default=0, help='cells per GPU (default: the BASELINE config size)')
    ap.add_argument('--dt', type=float, default=3600.0)
    ap.add_argument('--kernel', type=int, default=0, choices=[0, 1, 3])
    ap.add_argument('--mode', default='react', choices=['react', 'gi'],
                    help='react: operator-split RTReact (headline); gi: global-implicit auxvars + residual/Jacobian blocks (config 4)')
    ap.add_argument('--gi-dt', type=float, default=1800.0, dest='gi_dt')
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'],
                    help='weak (default, the driver\'s scaling run): cells per GPU fixed; strong: the single-GPU batch split over the ranks')
    ap.add_argument('--no-extra', action='store_true', help='headline only: skip the short runs of BASELINE configs 2 and 4')
    args = ap.parse_args()
    args.workload_given = any(a == '--workload' or a.startswith('--workload=') for a in sys.argv[1:])
    if args.mode == 'gi':
        if args.impl == 'reference':
            run_gi_reference(args)
        else:
            run_gi(args)
    elif args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
