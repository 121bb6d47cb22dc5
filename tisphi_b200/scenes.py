"""Scene dictionaries (tiSPHi scene-JSON schema) for the BASELINE.json configurations that are not shipped files."""
import copy

WATER = dict(matId=0, matType=1, density0=1000.0, viscosity=0.01, stiffness=500000.0, exponent=7.0, color=[50, 100, 200])

_COMMON = dict(GPUmemoryPercent=0.5, gravitation=[0.0, -9.81, 0.0], kappa=2.0, kh=1.5, boundary=2, kernel=1,
               kernelCorrection=0, colorTitle=7, colorGroup=0, showBdyPts=False, stepsPerRenderUpdate=10,
               pauseAtStart=False, stopEveryStep=0, stopAtStep=0, exitAtStep=0, stopAtTime=0, exitAtTime=0,
               exportEveryTime=0, exportEveryRender=0, exportFrame=False, exportVTK=False, exportCSV=False,
               kradius=1.0, givenMax=-1, givenMin=-1, fixMax=0, fixMin=0, comment="")


def dambreak3d(scale=1.0, precision="f32", **over):
    """BASELINE config C4: 3D WCSPH dambreak, d = 0.005 / scale.

    scale = 1: fluid block 1.6 x 1.0 x 0.8 m -> 320 x 200 x 160 = 10 240 000 fluid + 2 719 788 dummy = 12 959 788
    particles, 269 x 136 x 56 cells, dt = 2.4999999999999998e-05.  scale < 1 coarsens the lattice (same geometry)."""
    cfg = dict(_COMMON, is2D=False, particleRadius=0.0025 / scale, domainStart=[0.0, 0.0, 0.0], domainEnd=[4.0, 2.0, 0.8],
               simulationMethod=1, timeStepSizeMin=1e-6, timeIntegration=2, xsph=False, precision=precision)
    cfg.update(over)
    return {"Configuration": cfg, "Materials": [copy.deepcopy(WATER)],
            "Blocks": [dict(objectId=0, materialId=0, translation=[0.0, 0.0, 0.0], size=[1.6, 1.0, 0.8],
                            velocity=[0.0, 0.0, 0.0], rotationAxis=[0.0, 0.0, 1.0], rotationAngle=0.0)]}


def dambreak2d_small(precision="f64", **over):
    """The shrunken test1 dambreak used by smoke() and the fixtures (wc2d_small_lf)."""
    cfg = dict(_COMMON, is2D=True, particleRadius=0.01, domainStart=[0.0, 0.0, 0.0], domainEnd=[1.0, 0.6, 0.5],
               simulationMethod=1, timeStepSizeMin=1e-5, timeIntegration=2, xsph=False, precision=precision)
    cfg.update(over)
    return {"Configuration": cfg, "Materials": [copy.deepcopy(WATER)],
            "Blocks": [dict(objectId=0, materialId=0, translation=[0.0, 0.0, 0.0], size=[0.4, 0.3, 0.1],
                            velocity=[0.0, 0.0, 0.0], rotationAxis=[0.0, 0.0, 1.0], rotationAngle=0.0)]}
