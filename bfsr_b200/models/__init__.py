"""`models` registry surface of the reference (SRFlow-LP/code/models/__init__.py:24, LINF-LP/models/__init__.py)."""
from .models import make, register, models  # noqa: F401
from . import unet  # noqa: F401  (registers 'unet')
from .srflow import SRFlowModel, SRFlowNetEngine, define_Flow  # noqa: F401
from . import linf  # noqa: F401  (registers 'linf-patch')
from .linf import LINFEngine, LINFPriorEngine, batched_predict, batched_predict_log_p, build_inputs, lp_sr_mixed  # noqa: F401
