"""Model classes with the reference's names and ``spde_init`` dispatcher (``spdes/__init__.py:1-105``).

Every class is a configuration of :class:`spdepy_b200.spdes.base.SPDE2D`; the table below is the
reference's App. B parameter layout; the three separable classes live in ``separable.py``."""
from .base import SPDE2D

N9 = 9


def _cls(pyname, typename, **cfg):
    return type(pyname, (SPDE2D,), dict(name=typename, **cfg))


# spatial Whittle-Matern (whittle_matern2D.py, whittle_matern_anisotropic2D.py, whittle_matern_ha2D.py)
WhittleMatern2D = _cls("WhittleMatern2D", "whittle-matern-isotropic-2D", Hkind="iso", default_own=(-1, -1))
WhittleMaternAnisotropic2D = _cls("WhittleMaternAnisotropic2D", "whittle-matern-anisotropic-2D", Hkind="aniso",
                                  default_own=(-1, -1, 0.1, 0.1))
WhittleMaternHa2D = _cls("WhittleMaternHa2D", "whittle-matern-ha-2D", Hkind="ha", default_own=(-1, -1, 0.1, 0.1))
VarWhittleMaternAnisotropic2D = _cls("VarWhittleMaternAnisotropic2D", "var-whittle-matern-anisotropic-2D", kvar=True,
                                     Hkind="aniso", Hvar=True, default_own=([-1] * N9, [-1] * N9, [0.1] * N9, [0.1] * N9))
# (the reference's isotropic / half-angle var-Whittle-Matern classes assemble inline in logLike, var_whittle_matern2D.py:92-104,
#  var_whittle_matern_ha2D.py:126-140; here they get the common makeQ as well)
VarWhittleMatern2D = _cls("VarWhittleMatern2D", "var-whittle-matern-isotropic-2D", kvar=True, Hkind="iso", Hvar=True,
                          default_own=([-1] * N9, [-1] * N9))
VarWhittleMaternHa2D = _cls("VarWhittleMaternHa2D", "var-whittle-matern-ha-2D", kvar=True, Hkind="ha", Hvar=True,
                            default_own=([-1] * N9, [-1] * N9, [0.1] * N9, [0.1] * N9))
# advection-diffusion, constant coefficients (advection_*diffusion2D.py)
AdvectionDiffusion2D = _cls("AdvectionDiffusion2D", "advection-diffusion-2D", timed=True, Hkind="aniso", wkind="const",
                            aflav=1, default_own=(-1, -1, 0.01, 0.01, 0.01, 0.01, 0))
AdvectionIDiffusion2D = _cls("AdvectionIDiffusion2D", "advection-idiffusion-2D", timed=True, Hkind="iso", wkind="const",
                             aflav=1, default_own=(-1, -1, 0.1, 0.1, 0))
AdvectionHaDiffusion2D = _cls("AdvectionHaDiffusion2D", "advection-ha-diffusion-2D", timed=True, Hkind="ha", wkind="const",
                              aflav=1, default_own=(-1, -1, 0.01, 0.01, 0.01, 0.01, 0))
# spatially varying diffusion and/or advection
AdvectionVarDiffusion2D = _cls("AdvectionVarDiffusion2D", "advection-var-diffusion-2D", timed=True, kvar=True, Hkind="aniso",
                               Hvar=True, wkind="const", aflav=2,
                               default_own=([-1] * N9, [-1] * N9, [0.1] * N9, [0.1] * N9, 0.1, 0.1, 0))
AdvectionVarIDiffusion2D = _cls("AdvectionVarIDiffusion2D", "advection-var-idiffusion-2D", timed=True, kvar=True, Hkind="iso",
                                Hvar=True, wkind="const", aflav=2, default_own=([-1] * N9, [-1] * N9, 0.1, 0.1, 0))
VarAdvectionDiffusion2D = _cls("VarAdvectionDiffusion2D", "var-advection-diffusion-2D", timed=True, Hkind="aniso",
                               wkind="var", aflav=2, default_own=(-1, -1, 0.1, 0.1, [0.1] * 18, 0))
VarAdvectionIDiffusion2D = _cls("VarAdvectionIDiffusion2D", "var-advection-idiffusion-2D", timed=True, Hkind="iso",
                                wkind="var", aflav=2, default_own=(-1, -1, [0.01] * 18, 0))
AdvectionVarHaDiffusion2D = _cls("AdvectionVarHaDiffusion2D", "advection-var-ha-diffusion-2D", timed=True, kvar=True, Hkind="ha",
                                 Hvar=True, wkind="const", aflav=1,
                                 default_own=([-1] * N9, [-1] * N9, [0.1] * N9, [0.1] * N9, 0.1, 0.1, 0))
VarAdvectionHaDiffusion2D = _cls("VarAdvectionHaDiffusion2D", "var-advection-ha-diffusion-2D", timed=True, Hkind="ha",
                                 wkind="var", aflav=2, default_own=(-1, -1, 0.1, 0.1, [0.1] * 18, 0))
VarAdvectionVarHaDiffusion2D = _cls("VarAdvectionVarHaDiffusion2D", "var-advection-var-ha-diffusion-2D", timed=True,
                                    kvar=True, Hkind="ha", Hvar=True, wkind="var", aflav=2, divide=True,
                                    default_own=([-1] * N9, [-1] * N9, [0.1] * 18, [0.1] * 18, 0))
VarAdvectionVarDiffusion2D = _cls("VarAdvectionVarDiffusion2D", "var-advection-var-diffusion-2D", timed=True, kvar=True,
                                  Hkind="aniso", Hvar=True, wkind="var", aflav=2, divide=True,
                                  default_own=([-1] * N9, [-1] * N9, [0.1] * 18, [0.1] * 18, 0))
VarAdvectionVarIDiffusion2D = _cls("VarAdvectionVarIDiffusion2D", "var-advection-var-idiffusion-2D", timed=True, kvar=True,
                                   Hkind="iso", Hvar=True, wkind="var", aflav=2, divide=True,
                                   default_own=([-1] * N9, [-1] * N9, [0.1] * 18, 0))

# covariate-driven advection: w = lambda * ww with supplied face velocities ww (cov_advection_*2D.py)
CovAdvectionDiffusion2D = _cls("CovAdvectionDiffusion2D", "cov-advection-diffusion-2D", timed=True, Hkind="aniso", wkind="cov",
                               aflav=1, default_own=(-1, -1, 0.1, 0.1, 0.1, 0))
CovAdvectionIDiffusion2D = _cls("CovAdvectionIDiffusion2D", "cov-advection-idiffusion-2D", timed=True, Hkind="iso", wkind="cov",
                                aflav=1, default_own=(-1, -1, 0.1, 0))
CovAdvectionHaDiffusion2D = _cls("CovAdvectionHaDiffusion2D", "cov-advection-ha-diffusion-2D", timed=True, Hkind="ha",
                                 wkind="cov", aflav=1, default_own=(-1, -1, 0.01, 0.01, 0.01, 0))
CovAdvectionVarDiffusion2D = _cls("CovAdvectionVarDiffusion2D", "cov-advection-var-diffusion-2D", timed=True, kvar=True,
                                  Hkind="aniso", Hvar=True, wkind="cov", aflav=2,
                                  default_own=([-1] * N9, [-1] * N9, [0.1] * N9, [0.1] * N9, 0.1, 0))
CovAdvectionVarIDiffusion2D = _cls("CovAdvectionVarIDiffusion2D", "cov-advection-var-idiffusion-2D", timed=True, kvar=True,
                                   Hkind="iso", Hvar=True, wkind="cov", aflav=2, default_own=([-1] * N9, [-1] * N9, 0.1, 0))
CovAdvectionVarHaDiffusion2D = _cls("CovAdvectionVarHaDiffusion2D", "cov-advection-var-ha-diffusion-2D", timed=True, kvar=True,
                                    Hkind="ha", Hvar=True, wkind="cov", aflav=2,
                                    default_own=([-1] * N9, [-1] * N9, [0.1] * N9, [0.1] * N9, 0.1, 0))

_TABLE = {
    # model id -> (ha class, anisotropic class, isotropic class)
    ("whittle-matern", 1): (WhittleMaternHa2D, WhittleMaternAnisotropic2D, WhittleMatern2D),
    ("var-whittle-matern", -1): (VarWhittleMaternHa2D, VarWhittleMaternAnisotropic2D, VarWhittleMatern2D),
    ("advection-diffusion", 2): (AdvectionHaDiffusion2D, AdvectionDiffusion2D, AdvectionIDiffusion2D),
    ("advection-var-diffusion", 3): (AdvectionVarHaDiffusion2D, AdvectionVarDiffusion2D, AdvectionVarIDiffusion2D),
    ("cov-advection-diffusion", 4): (CovAdvectionHaDiffusion2D, CovAdvectionDiffusion2D, CovAdvectionIDiffusion2D),
    ("cov-advection-var-diffusion", 5): (CovAdvectionVarHaDiffusion2D, CovAdvectionVarDiffusion2D, CovAdvectionVarIDiffusion2D),
    ("var-advection-diffusion", 6): (VarAdvectionHaDiffusion2D, VarAdvectionDiffusion2D, VarAdvectionIDiffusion2D),
    ("var-advection-var-diffusion", 7): (VarAdvectionVarHaDiffusion2D, VarAdvectionVarDiffusion2D, VarAdvectionVarIDiffusion2D),
}
_NEXT = {"seperable-spatial-temporal": 8}


def spde_init(model, grid, parameters=None, ani=True, ha=True, bc=3, mod0=None):
    if grid.sdim != 2:
        raise NotImplementedError("Model not implemented for 3D grids")
    for (name, num), (c_ha, c_ani, c_iso) in _TABLE.items():
        if model == name or model == num:
            cls = c_ha if ha else (c_ani if ani else c_iso)
            if cls is None:
                raise NotImplementedError("%s with ha=%s, anisotropic=%s is not wired to the B200 path yet "
                                          "(SURVEY.md section 8f)" % (name, ha, ani))
            if cls.timed:
                return cls(par=parameters, grid=grid, bc=bc, mod0=mod0)
            return cls(par=parameters, grid=grid, bc=bc)
    if model in _NEXT or model in _NEXT.values():
        from . import separable as sep
        cls = (sep.SeperableSpatialTemporalHa2D if ha else
               (sep.SeperableSpatialTemporal2D if ani else sep.SeperableSpatialTemporalIDiffusion2D))
        return cls(par=parameters, grid=grid, bc=bc)
    raise AssertionError("Model not implemented")
