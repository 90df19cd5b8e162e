"""svo_b200 -- Python host binding of libsvo_b200.so (B200-native SVO raycaster).

The directory name carries the reference's name (sparse-voxel-octree-raycasting_b200) and is not a valid
Python identifier; load it with ``importlib`` (see ``__graft_entry__.load_package()``) under the module
name ``svo_b200``.

This package is plumbing over the C ABI in include/svo_b200.h (ctypes; numpy arrays for host buffers).
It never falls back to a CPU implementation: a missing or unbuilt CUDA library raises ImportError here,
and a missing GPU raises at ``ocl_init()``.
"""
from .ocl import (Device, Mem, ocl_init, ocl_exit, ocl_get_kernel, ocl_malloc, ocl_copy_to_host, ocl_copy_to_device,
                  ocl_begin, ocl_param, ocl_end, ocl_begin_all_kernels, ocl_end_all_kernels, ocl_memcpy, ocl_memset,
                  ocl_round_up, lib, LIB_PATH, FrameParams, launch_count, set_octree_depth)
from . import raycast, scene, bands  # noqa: F401

__all__ = ["Device", "Mem", "ocl_init", "ocl_exit", "ocl_get_kernel", "ocl_malloc", "ocl_copy_to_host",
           "ocl_copy_to_device", "ocl_begin", "ocl_param", "ocl_end", "ocl_begin_all_kernels", "ocl_end_all_kernels",
           "ocl_memcpy", "ocl_memset", "ocl_round_up", "lib", "LIB_PATH", "FrameParams", "launch_count",
           "set_octree_depth", "raycast", "scene", "bands"]
