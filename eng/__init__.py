"""Module path of the reference (``from eng.simulation import Simulation, SimConfiger``, run_simulation.py:3-4 of
Rabmelon/tiSPHi; SURVEY 8b: the ``eng.*`` paths are part of the surface to preserve).  Every module here re-exports
``tisphi_b200.eng.<same name>``: with this repository's root on ``sys.path`` the reference's own entry script runs
against the CUDA engine after deleting its two Taichi lines (``import taichi as ti`` and ``ti.init(...)``)."""
