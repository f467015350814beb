from .meters import AverageMeter, accuracy, Timer, DeviceTimer, Metric  # noqa: F401
from .logging import TableLogger, TSVLogger, JsonlLogger, ScalarWriter  # noqa: F401
from .checkpoint import save_checkpoint, load_state, make_state, find_latest  # noqa: F401
from .profiling import nvtx_range, trace, kernel_table  # noqa: F401
