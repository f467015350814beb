from cpd_b200.parallel.api import (DistModule, sum_gradients, normal_sum_gradients, kahan_sum_gradients,  # noqa: F401
                                   broadcast_params)
from cpd_b200.parallel.dist import dist_init  # noqa: F401
