from .delayed_gpu_accelerator import _GPUAccelerator

__all__ = ["_GPUAccelerator"]
