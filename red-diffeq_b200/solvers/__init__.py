from .pde import FWIForward
from .sharding import ShardedFWIForward, plan_partition, split_range

__all__ = ["FWIForward", "ShardedFWIForward", "plan_partition", "split_range"]
