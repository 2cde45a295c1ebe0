"""B200-native (sm_100a) replacement for the TensorRT engine path of
yester31/Monocular_Depth_Estimation_TRT: `get_engine` / `allocate_buffers` / `do_inference` /
`free_buffers` / `StageTimer` with the reference's call shapes, over libmde_b200.so.

Importing the package does not load the shared library; the first engine or kernel call does, and
fails loudly if it has not been built (there is no fallback path).
"""
__all__ = ["common", "common_runtime", "engine", "weights", "build"]
__version__ = "0.1.0"
