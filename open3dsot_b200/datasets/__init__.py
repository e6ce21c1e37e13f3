from .synthetic import synthetic_siamese_batch, synthetic_motion_batch  # noqa: F401
