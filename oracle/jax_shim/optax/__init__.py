"""Import-only stub: gpfy.model / gpfy.training name these types at import time."""
GradientTransformation = object
OptState = object
Params = object


def _unavailable(*_a, **_k):
    raise NotImplementedError("optax is not available in the jax_shim")


chain = masked = set_to_zero = apply_updates = adam = _unavailable
