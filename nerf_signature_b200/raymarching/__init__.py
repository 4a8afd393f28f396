from .raymarching import *  # noqa: F401,F403  (same surface as the reference's raymarching/__init__.py)
from .raymarching import (near_far_from_aabb, sph_from_ray, morton3D, morton3D_invert, packbits,
                          march_rays_train, composite_rays_train, march_rays, composite_rays)
