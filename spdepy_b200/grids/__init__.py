"""``grid(x, y, z=None, t=None, extend=None, Nbs=3)`` -- same constructor as the reference
(``grids/__init__.py:3-49``).  3-D meshes are out of scope (no reference model uses them,
``spdes/__init__.py:104-105``)."""
from .regular_mesh import GridS, GridST


def grid(x, y, z=None, t=None, extend=None, Nbs=3):
    if z is not None:
        raise NotImplementedError("3-D meshes are outside the hot path (no reference model uses them)")
    mesh = GridS() if t is None else GridST()
    if t is None:
        mesh.setGrid(x=x, y=y, extend=extend, Nbs=Nbs)
    else:
        mesh.setGrid(x=x, y=y, t=t, extend=extend, Nbs=Nbs)
    return mesh
