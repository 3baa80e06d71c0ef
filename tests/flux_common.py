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


def boundary_connections(nx, ny, nz, naq, seed=21, ghost_layers=0, g2l=None):
    """Boundary faces of the structured block in the reference's order of boundary conditions (west, east, south, north, bottom,
    top regions, each a coupler of patch%boundary_condition_list): id_dn = ghosted id of the local cell behind the face.  Edge and
    corner cells appear in two / three faces.  Returns dict(id_dn, area, velocity, disp, bc_type)."""
    g = ghost_layers
    gx, gy, gz = nx + 2 * g, ny + 2 * g, nz + 2 * g
    idx = np.arange(gx * gy * gz, dtype=np.int64).reshape(gz, gy, gx)
    inner = idx[g:gz - g, g:gy - g, g:gx - g]
    faces = [inner[:, :, 0], inner[:, :, -1], inner[:, 0, :], inner[:, -1, :], inner[0, :, :], inner[-1, :, :]]
    id_dn = np.concatenate([f.ravel() for f in faces]).astype(np.int32)
    bc_type = np.concatenate([np.full(f.size, t, dtype=np.int32) for f, t in zip(faces, [1, 3, 4, 1, 3, 4])])   # DIRICHLET / DIRICHLET_ZERO_GRADIENT / ZERO_GRADIENT
    rng = np.random.default_rng(seed)
    nb = id_dn.size
    out = {'id_dn': np.ascontiguousarray(id_dn), 'id_up': np.ascontiguousarray(id_dn), 'bc_type': bc_type,   # id_up: only sizes Oracle.flux_coefs
           'area': rng.uniform(0.5, 2.0, nb),
           'velocity': rng.normal(0.0, 1.0e-6, nb), 'disp': rng.uniform(1.0e-9, 1.0e-7, (nb, naq))}
    out['velocity'][::13] = 0.0
    return out


def source_sinks(ncells_local_ghosted_ids, naq, nss=9, seed=23):
    """A few wells: injection (qsrc > 0), extraction (qsrc < 0), one EQUILIBRIUM_SS (12) and one MASS_RATE_SS (7) coupler; two wells
    share a cell.  Returns dict(id_dn, qsrc, ss_type)."""
    rng = np.random.default_rng(seed)
    ids = rng.choice(ncells_local_ghosted_ids, nss, replace=False).astype(np.int32)
    ids[1] = ids[0]
    qsrc = rng.normal(0.0, 1.0e-5, nss)
    qsrc[2] = 0.0
    ss_type = np.zeros(nss, dtype=np.int32)
    ss_type[3] = 12
    ss_type[4] = 7
    return {'id_dn': np.ascontiguousarray(ids), 'qsrc': qsrc, 'ss_type': ss_type}
