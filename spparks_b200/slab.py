"""Slab decomposition bookkeeping shared by bench.py and the tests (host side of SURVEY §8e).

The last lattice dimension (z in 3-D, y in 2-D) is cut into `nranks` equal slabs, the analogue of
`processors 1 1 N` / `processors 1 N 1` (src/domain.cpp:204-328).  Each slab keeps one ghost plane
per side (delpropensity = 1) refreshed from its ring neighbours before every colour set, as
CommLattice::all() does (src/comm_lattice.cpp:137-143)."""

DIM = {"sq/4n": 2, "sq/8n": 2, "tri": 2, "sc/6n": 3, "sc/26n": 3}
NBASIS = {"sq/4n": 1, "sq/8n": 1, "tri": 2, "sc/6n": 1, "sc/26n": 1}


def slab_extent(lattice, n):
    """Global extent of the slab dimension."""
    return n[2] if DIM[lattice] == 3 else n[1]


def slab_bounds(lattice, n, rank, nranks):
    """[lo, hi) cell range of the slab dimension owned by `rank`; thickness must be even so that
    the colour parity of local and global indices agree (src/lattice.cpp:234-283)."""
    ext = slab_extent(lattice, n)
    if ext % nranks:
        raise ValueError("slab dimension %d does not split into %d slabs" % (ext, nranks))
    t = ext // nranks
    if nranks > 1 and (t % 2 or t < 2):
        raise ValueError("slab thickness %d must be even and >= 2" % t)
    return rank * t, (rank + 1) * t


def plane_sites(lattice, n):
    """Sites in one plane (3-D) / row (2-D) of the slab dimension."""
    return n[0] * n[1] * NBASIS[lattice] if DIM[lattice] == 3 else n[0] * NBASIS[lattice]


def site_range(lattice, n, rank, nranks):
    """[first, last) reference site indices owned by `rank` (sites are ordered slab-dimension slowest)."""
    lo, hi = slab_bounds(lattice, n, rank, nranks)
    p = plane_sites(lattice, n)
    return lo * p, hi * p


def ring(rank, nranks):
    """(lower, upper) ring neighbours: my first plane is the upper ghost of `lower`, my last plane
    the lower ghost of `upper`."""
    return (rank - 1) % nranks, (rank + 1) % nranks


def slab_box(lattice, n, rank, nranks):
    """The slab as a sector-style box {ilo,ihi,jlo,jhi,klo,khi} of global cells."""
    lo, hi = slab_bounds(lattice, n, rank, nranks)
    if DIM[lattice] == 3:
        return [0, n[0], 0, n[1], lo, hi]
    return [0, n[0], lo, hi, 0, 1]
