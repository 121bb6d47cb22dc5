"""Re-export of tisphi_b200.eng.ui_sim under the reference's module path (see eng/__init__.py)."""
from tisphi_b200.eng.ui_sim import *  # noqa: F401,F403
from tisphi_b200.eng import ui_sim as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
