"""Import-path drop-in: ``from cpp_raytracer.raytracer_cpp import RayTracer, Scene, Sphere, Material,
Vector3, Camera`` (reference interaction.py:13, gui.py:12, run.py:54) resolves to the B200 path.
Like the reference's, this directory is a namespace package (no __init__.py); put the repo root on
sys.path (or copy this one file next to the reference's interaction.py as cpp_raytracer/raytracer_cpp.py)."""
from pgr_raytracing_project_b200.raytracer_cpp import (  # noqa: F401
    Camera, DebugInfo, Material, Ray, RayTracer, Scene, Sphere, Vector3)
