from . import function_call_parser as _fcp  # noqa: F401
from . import more_detectors as _more  # noqa: F401  (registers the additional formats in DETECTORS)
