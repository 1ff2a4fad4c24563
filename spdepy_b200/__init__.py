"""spdepy_b200 -- B200-native precision-and-likelihood hot path behind spdepy's model-class API.

    import spdepy_b200 as sp
    mod = sp.model(grid=sp.grid(x, y, t), spde="advection-diffusion", ha=False, bc=3, mod0=...)

Same factory surface as the reference (``src/spdepy/__init__.py:13-52``)."""
from .grids import grid
from .model import Model
from .optim import Optimize as optim
from .spdes import spde_init

__version__ = "0.1.0"


def model(**kwargs) -> Model:
    assert kwargs.get("grid") is not None and "grid" in kwargs.get("grid").type, "Grid is not defined"
    ha = True if type(kwargs.get("ha")) is not bool else kwargs.get("ha")
    bc = 3 if type(kwargs.get("bc")) is not int else kwargs.get("bc")
    ani = True if type(kwargs.get("anisotropic")) is not bool else kwargs.get("anisotropic")
    mesh = kwargs.get("grid")
    if mesh.type == "gridST" and kwargs.get("mod0") is None:
        mesh0 = grid(x=mesh.x, y=mesh.y, extend=mesh.Ne or None)
        mod0 = spde_init(model="whittle-matern", grid=mesh0, ani=ani, ha=ha, bc=bc)
    elif mesh.type == "gridST":
        mod0 = kwargs.get("mod0").mod
    else:
        mod0 = None
    return Model(spde=kwargs.get("spde"), grid=mesh, parameters=kwargs.get("parameters"), ani=ani, ha=ha, bc=bc, mod0=mod0)
