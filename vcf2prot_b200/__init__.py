"""vcf2prot_b200 -- B200-native sequence-generation engine behind vcf2prot's `Engine::GPU`.

Only the path SURVEY.md section 8 names lives here: the C ABI + CUDA kernels (csrc/, libv2p_engine.so) and a
thin host mirror of the reference's Task / GIR / Engine interface (engine.py) used by tests and bench.py.
"""
from .engine import Engine, EngineError, GIR, GpuEngine, Task, pack_tasks  # noqa: F401

__all__ = ["Engine", "EngineError", "GIR", "GpuEngine", "Task", "pack_tasks"]
