from .dist import dist_init, discover_topology, detect_launcher, get_rank, get_world_size, is_distributed, Topology, simple_group_split  # noqa: F401
from .arena import ArenaLayout, ModelArena, ChunkTable  # noqa: F401
from .reducer import GradReducer  # noqa: F401
from .emulate import EmulateNode  # noqa: F401
from .api import (sum_gradients, normal_sum_gradients, kahan_sum_gradients, broadcast_params, DistModule)  # noqa: F401
from .hooks import APSOptimizerHook  # noqa: F401
