"""Process-wide execution switches (read by the modules; no global state inside the native library)."""
import contextlib
import os

_FUSED = os.environ.get("O3D_FUSED", "1") != "0"


# tcgen05 3xTF32 GEMM core for the point-wise layers: bit 0 = forward + dgrad, bit 1 = wgrad  (0 = exact-fp32 CUDA cores)
_TC = int(os.environ.get("O3D_TC", "3"))


# Lifted first layer of grouped stacks (include/o3d_b200.h o3d_lift_t): 1 = the grouped tensor is never built (default),
# 0 = materialising ball-query+group kernel followed by a GEMM over the grouped rows (round-1 path, kept as a cross-check)
_LIFT = os.environ.get("O3D_LIFT", "1") != "0"


def lift_enabled() -> bool:
    return _LIFT


def set_lift(flag: bool) -> None:
    global _LIFT
    _LIFT = bool(flag)


def tc_enabled() -> bool:
    return _TC != 0


def tc_level() -> int:
    return _TC


def set_tc(level) -> None:
    global _TC
    _TC = int(level)


def fused_enabled() -> bool:
    return _FUSED


def set_fused(flag: bool) -> None:
    global _FUSED
    _FUSED = bool(flag)


@contextlib.contextmanager
def composed_mode():
    """Run modules as the reference's op-by-op composition over the nine `_ext` kernels (cross-check path)."""
    global _FUSED
    old = _FUSED
    _FUSED = False
    try:
        yield
    finally:
        _FUSED = old
