from .distributed_upscale import UltimateSDUpscaleDistributed
from .collector import DistributedCollectorNode

NODE_CLASS_MAPPINGS = {
    "UltimateSDUpscaleDistributed": UltimateSDUpscaleDistributed,
    "DistributedCollector": DistributedCollectorNode,
}
NODE_DISPLAY_NAME_MAPPINGS = {
    "UltimateSDUpscaleDistributed": "Ultimate SD Upscale Distributed (No Upscale)",
    "DistributedCollector": "Distributed Collector",
}
