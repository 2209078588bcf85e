"""The "_gpu" accelerator: claims to be available on a GPU-less driver and does NOT bind a CUDA
device at ``setup_environment`` — the worker binds it later, once its root device is known
(ray_lightning/accelerators/delayed_gpu_accelerator.py:22-60; binding happens in
launchers/ray_launcher.py via util.set_cuda_device_if_used)."""
from typing import Dict, List

import torch


class _GPUAccelerator:
    def setup_environment(self, root_device: torch.device) -> None:
        pass  # deliberately no torch.cuda.set_device here

    @staticmethod
    def get_parallel_devices(devices: List[int]) -> List[torch.device]:
        return [torch.device("cuda", i) for i in devices] if devices else []

    @staticmethod
    def is_available() -> bool:
        return True  # the driver may have no GPU; the workers do

    @classmethod
    def register_accelerators(cls, accelerator_registry: Dict) -> None:
        accelerator_registry.register("_gpu", cls, description=cls.__name__)
