"""Re-export of tisphi_b200.eng.particle_system under the reference's module path (see eng/__init__.py)."""
from tisphi_b200.eng.particle_system import *  # noqa: F401,F403
from tisphi_b200.eng import particle_system as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
