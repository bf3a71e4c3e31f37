from .distributed_upscale import UltimateSDUpscaleDistributed
from .collector import DistributedCollectorNode
from .utilities import ImageBatchDivider

NODE_CLASS_MAPPINGS = {
    "UltimateSDUpscaleDistributed": UltimateSDUpscaleDistributed,
    "DistributedCollector": DistributedCollectorNode,
    "ImageBatchDivider": ImageBatchDivider,
}
NODE_DISPLAY_NAME_MAPPINGS = {
    "UltimateSDUpscaleDistributed": "Ultimate SD Upscale Distributed (No Upscale)",
    "DistributedCollector": "Distributed Collector",
    "ImageBatchDivider": "Image Batch Divider",
}
