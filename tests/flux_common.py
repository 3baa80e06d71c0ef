"""Synthetic connection lists for the flux-side tests: a structured nx*ny*nz block in the reference's
connection order (all x connections, then y, then z — grid_structured.F90 StructGridComputeInternConnect), optional
ghost layer (cells that are ghosted but not local) and inactive cells."""
import numpy as np


def structured_connections(nx, ny, nz, naq, seed=3, ghost_layers=0, inactive_fraction=0.0):
    """Returns (conn dict, nghosted, nlocal, active).  With ghost_layers = g the ghosted block is (nx+2g)(ny+2g)(nz+2g) and the
    local cells are its interior, numbered in natural order (as grid%nG2L does)."""
    g = ghost_layers
    gx, gy, gz = nx + 2 * g, ny + 2 * g, nz + 2 * g
    idx = np.arange(gx * gy * gz, dtype=np.int64).reshape(gz, gy, gx)
    ups, dns = [], []
    ups.append(idx[:, :, :-1].ravel()); dns.append(idx[:, :, 1:].ravel())
    ups.append(idx[:, :-1, :].ravel()); dns.append(idx[:, 1:, :].ravel())
    ups.append(idx[:-1, :, :].ravel()); dns.append(idx[1:, :, :].ravel())
    id_up = np.concatenate(ups).astype(np.int32)
    id_dn = np.concatenate(dns).astype(np.int32)
    nghosted = gx * gy * gz
    g2l = None
    nlocal = nghosted
    if g > 0:
        g2l = np.full(nghosted, -1, dtype=np.int32)
        interior = idx[g:gz - g, g:gy - g, g:gx - g].ravel()
        g2l[interior] = np.arange(interior.size, dtype=np.int32)
        nlocal = interior.size
        # the reference keeps only connections with at least one local side
        keep = (g2l[id_up] >= 0) | (g2l[id_dn] >= 0)
        id_up, id_dn = id_up[keep], id_dn[keep]
    rng = np.random.default_rng(seed)
    nconn = id_up.size
    conn = {
        'id_up': np.ascontiguousarray(id_up), 'id_dn': np.ascontiguousarray(id_dn), 'g2l': g2l,
        'area': rng.uniform(0.5, 2.0, nconn),
        'velocity': rng.normal(0.0, 1.0e-6, nconn),            # Darcy flux, both signs (both upwinding branches)
        'disp': rng.uniform(1.0e-9, 1.0e-7, (nconn, naq)),      # harmonic dispersion / distance per component
        'fraction_upwind': rng.uniform(0.3, 0.7, nconn),
    }
    conn['velocity'][::17] = 0.0
    active = np.ones(nghosted, dtype=np.uint8)
    if inactive_fraction > 0:
        active[rng.random(nghosted) < inactive_fraction] = 0
    return conn, nghosted, nlocal, active


def random_connections(ncells, nconn, naq, seed=9, nghost=0):
    """Unstructured connectivity: nconn random pairs (repeats of a pair allowed, as two faces between the same cells), a few hub
    cells with many connections (rows far longer than a structured grid's six), the last nghost ghosted cells not local."""
    rng = np.random.default_rng(seed)
    up = rng.integers(0, ncells, nconn)
    dn = rng.integers(0, ncells, nconn)
    hubs = rng.choice(ncells, 3, replace=False)
    up[::7] = hubs[0]
    dn[3::11] = hubs[1]
    same = up == dn
    dn[same] = (dn[same] + 1) % ncells
    g2l = None
    nlocal = ncells
    if nghost:
        g2l = np.arange(ncells, dtype=np.int32)
        g2l[ncells - nghost:] = -1
        nlocal = ncells - nghost
    conn = {
        'id_up': up.astype(np.int32), 'id_dn': dn.astype(np.int32), 'g2l': g2l,
        'area': rng.uniform(0.5, 2.0, nconn), 'velocity': rng.normal(0.0, 1.0e-6, nconn),
        'disp': rng.uniform(1.0e-9, 1.0e-7, (nconn, naq)), 'fraction_upwind': rng.uniform(0.3, 0.7, nconn),
    }
    return conn, ncells, nlocal, np.ones(ncells, dtype=np.uint8)
