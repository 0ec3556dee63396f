"""B200-native (sm_100a CUDA) render hot path of Samuel-2000/PGR-Raytracing-Project.

* ``raytracer_cpp``  -- drop-in for the reference's ``cpp_raytracer.raytracer_cpp`` module
* ``context``        -- RenderContext: the C-ABI library (include/b200rt.h) + torch device buffers
* ``scenes``         -- synthetic inputs of the benchmark configurations (plain numpy)
* ``multigpu``       -- tile-sharded rendering across ranks (torch.distributed, NCCL gather)
* ``build``          -- nvcc build of libb200rt.so for sm_100a

Importing this package does not load the CUDA library; ``context``/``raytracer_cpp`` do, and they
fail loudly when it is missing or no CUDA device is present (there is no CPU fallback).
"""
__version__ = "0.1.0"
