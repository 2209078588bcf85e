from .ray_launcher import RayLauncher
from .ray_horovod_launcher import RayHorovodLauncher

__all__ = ["RayLauncher", "RayHorovodLauncher"]
