from .fea_model import FEAModel  # noqa: F401
from .state_model import StateModel, StateOperation  # noqa: F401
from .output_model import OutputModel, OutputOperation, OutputFieldModel, OutputFieldOperation  # noqa: F401
from ._csdl_compat import Simulator  # noqa: F401
