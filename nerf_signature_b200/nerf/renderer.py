"""Clean-pipeline renderer (reference nerf/renderer.py).  The reference keeps two copies of the
renderer that differ only by the `message` argument (a 15-line diff); here one class serves both,
`message` defaulting to None."""
from .renderer_wtmk import NeRFRenderer, sample_pdf, custom_meshgrid  # noqa: F401
