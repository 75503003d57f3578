"""x-slab domain decomposition and grid set-up (host logic, pure Python).

Restates mpi_routines.F90:312-337 (cell ranges per rank with nprocy = 1), setup.F90:164-206
(global grid), utilities.f90:343-372 (setup_grid_x) and setup.F90:629-646 (set_dt).  These
stay on the host in the reference too; the device library receives the resulting scalars
through `cylgpu_config`.
"""
import math
from dataclasses import dataclass

from .constants import C_LIGHT


def slab_bounds(nx_global, nranks):
    """[(cell_x_min, cell_x_max)] per rank, 1-based inclusive (mpi_routines.F90:312-337)."""
    nx0 = nx_global // nranks
    if nx0 * nranks != nx_global:
        nxp = (nx0 + 1) * nranks - nx_global
    else:
        nxp = nranks
    out = []
    for idim in range(1, nranks + 1):
        if idim <= nxp:
            lo, hi = (idim - 1) * nx0 + 1, idim * nx0
        else:
            lo = nxp * nx0 + (idim - nxp - 1) * (nx0 + 1) + 1
            hi = nxp * nx0 + (idim - nxp) * (nx0 + 1)
        out.append((lo, hi))
    return out


@dataclass
class SlabGrid:
    """Grid scalars of one rank; `shift()` advances the moving window by one cell."""
    nx_global: int
    ny_global: int
    nranks: int
    rank: int
    x_min: float
    x_max: float
    y_max: float
    dt_multiplier: float = 0.95
    bounds: list = None        # [(cell_x_min, cell_x_max)] of every rank after a re-balance; None: the even split

    def __post_init__(self):
        self.length_x = self.x_max - self.x_min
        self.dx = self.length_x / float(self.nx_global)
        self.dy = (self.y_max - 0.0) / float(self.ny_global)
        self.xb_min = self.x_min                      # cell-edge origin
        self.x_grid_min = self.x_min + self.dx / 2.0  # cell-centre origin
        self.y_grid_min_local = 0.0 + self.dy / 2.0
        self.cell_x_min, self.cell_x_max = (self.bounds or slab_bounds(self.nx_global, self.nranks))[self.rank]
        self.nx = self.cell_x_max - self.cell_x_min + 1
        self.ny = self.ny_global
        self.x_min_boundary = self.rank == 0
        self.x_max_boundary = self.rank == self.nranks - 1
        dt = 0.9 * min(self.dx, self.dy) / math.sqrt(2.0) / C_LIGHT   # setup.F90:639
        self.dt = self.dt_multiplier * dt                             # setup.F90:646
        self.setup_grid_x()

    @classmethod
    def like(cls, other, bounds, rank):
        """the grid of `rank` for new slab bounds, every scalar of the (possibly shifted) window taken over as it is:
        x_grid_min and the box edges are running sums of dx (window.F90:76-86), not functions of each other"""
        g = cls(other.nx_global, other.ny_global, other.nranks, rank, other.x_min, other.x_max, other.y_max,
                other.dt_multiplier, bounds=[tuple(b) for b in bounds])
        for name in ("length_x", "dx", "dy", "xb_min", "x_grid_min", "y_grid_min_local", "x_min", "x_max", "dt"):
            setattr(g, name, getattr(other, name))
        g.setup_grid_x()
        return g

    def setup_grid_x(self):   # utilities.f90:343-372 with cpml offsets = 0
        self.x_grid_min_local = self.x_grid_min + float(self.cell_x_min - 1) * self.dx
        self.x_grid_max_local = self.x_grid_min + float(self.cell_x_max - 1) * self.dx
        self.x_min_local = self.x_grid_min_local + (0 - 0.5) * self.dx
        self.x_max_local = self.x_grid_max_local - (0 - 0.5) * self.dx

    def shift(self):          # window.F90:62-94 grid part
        self.x_grid_min = self.x_grid_min + self.dx
        self.xb_min = self.xb_min + self.dx
        self.x_min = self.xb_min
        self.x_max = self.xb_min + float(self.nx_global) * self.dx
        self.setup_grid_x()

    @property
    def x_grid_max(self):
        return self.x_grid_min + float(self.nx_global - 1) * self.dx
