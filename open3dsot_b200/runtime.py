"""Process-wide execution switches (read by the modules; no global state inside the native library)."""
import contextlib
import os

_FUSED = os.environ.get("O3D_FUSED", "1") != "0"


# tcgen05 3xTF32 GEMM core for the point-wise layers: bit 0 = forward + dgrad, bit 1 = wgrad  (0 = exact-fp32 CUDA cores)
_TC = int(os.environ.get("O3D_TC", "3"))


# Lifted first layer of grouped stacks (include/o3d_b200.h o3d_lift_t): 1 = the grouped tensor is never built (default),
# 0 = materialising ball-query+group kernel followed by a GEMM over the grouped rows (round-1 path, kept as a cross-check)
_LIFT = os.environ.get("O3D_LIFT", "1") != "0"


LIFT_FILTER = None      # diagnostics: callable(site: str, info: dict) -> bool restricting which stacks are lifted


def lift_enabled(site=None, info=None) -> bool:
    if not _LIFT:
        return False
    if LIFT_FILTER is not None and site is not None:
        return bool(LIFT_FILTER(site, info or {}))
    return True


def set_lift(flag: bool) -> None:
    global _LIFT
    _LIFT = bool(flag)


# Inference with static weights (the tracking loop): eval-mode stacks pack their weights and fold their running BatchNorm
# statistics ONCE (o3d_stack_prepare) instead of per call.  Off by default — a cached block goes stale when the weights change
# (the engine's fused Adam updates parameters in place without bumping tensor versions): turn it on only around inference,
# and call fused.clear_prepared() after loading new weights.
_SA_FUSED = os.environ.get("O3D_SA_FUSED", "1") != "0"


def sa_fused_enabled() -> bool:
    """eval-mode forward passes (no autograd) run every set-abstraction layer as ONE kernel (csrc/sa_fused.cu)"""
    return _SA_FUSED


def set_sa_fused(flag: bool) -> None:
    global _SA_FUSED
    _SA_FUSED = bool(flag)


_BRANCH_OVERLAP = os.environ.get("O3D_BRANCH_OVERLAP", "1") != "0"


def branch_overlap_enabled() -> bool:
    """inference: the template branch runs on a side stream next to the search branch (fused.run_ahead)"""
    return _BRANCH_OVERLAP


def set_branch_overlap(flag: bool) -> None:
    global _BRANCH_OVERLAP
    _BRANCH_OVERLAP = bool(flag)


_STATIC_WEIGHTS = False


def static_weights() -> bool:
    return _STATIC_WEIGHTS


def set_static_weights(flag: bool) -> None:
    global _STATIC_WEIGHTS
    _STATIC_WEIGHTS = bool(flag)


@contextlib.contextmanager
def static_weights_scope():
    """`with runtime.static_weights_scope():` — inference code whose weights do not change while it runs."""
    global _STATIC_WEIGHTS
    old = _STATIC_WEIGHTS
    _STATIC_WEIGHTS = True
    try:
        yield
    finally:
        _STATIC_WEIGHTS = old


# In-place accumulation of parameter gradients: a stack's backward ADDS its weight / bias / BatchNorm gradients straight into the
# parameters' existing `.grad` buffers (and returns no gradient for them) instead of materialising them and letting autograd's
# AccumulateGrad issue one elementwise add per parameter — ~130 launches per BAT step.  Only valid for `loss.backward()` onto
# pre-allocated `.grad` buffers (the engine's flat bucket); `torch.autograd.grad` callers must leave it off (default).
_GRAD_INPLACE = False


def grad_inplace() -> bool:
    return _GRAD_INPLACE


@contextlib.contextmanager
def grad_inplace_scope():
    global _GRAD_INPLACE
    old = _GRAD_INPLACE
    _GRAD_INPLACE = True
    try:
        yield
    finally:
        _GRAD_INPLACE = old


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


# ---- discrete-choice hook (parity tests only) -------------------------------------------------------------------
# The forward pass takes three kinds of data-dependent DISCRETE decisions on computed values: the ball query of every
# set-abstraction layer (on input coordinates in the backbone, on VOTED coordinates in the RPN), and BoxAwareXCorr's top-k
# over predicted box clouds.  A neighbour that sits within fp32 round-off of the radius / of the k-th distance can fall the
# other way on the GPU than in the CPU oracle, which changes downstream floats by O(1e-3) without any kernel being wrong.
# Parity tests install a hook that (a) records the product's own choice and (b) may substitute the oracle's, so that the
# float path is compared at 1e-4 with identical discrete choices.  hook(kind, info, compute) -> int32 tensor; `compute()`
# evaluates the product's choice.  None (the default) = no hook: the product path is unchanged.
CHOICE_HOOK = None


def choose(kind, info, compute):
    if CHOICE_HOOK is None:
        return compute()
    return CHOICE_HOOK(kind, info, compute)
